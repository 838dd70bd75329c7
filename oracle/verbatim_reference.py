"""ORACLE tooling (container-only): import the reference's own modules VERBATIM.

`/root/reference/main/{actionspace,gameimage,game,mcts}.py` import `chess`,
`config` (which imports TensorFlow, config.py:1) and `model` (TensorFlow/Keras,
model.py:1-2) at top level; none of those are installed here.  This helper
registers throw-away stand-ins in `sys.modules` so the four files can be
imported unmodified, straight from the read-only reference tree:

* `chess`  -> oracle.pychess_shim (from-scratch python-chess subset)
* `config` -> a `Config` with the values of config.py:3-42 minus the TF enum
* `model`  -> `predict_fn(network, image)` that calls `network(image[None])`
              eagerly and unpacks it exactly like model.py:73-80

It is used ONLY to validate the restated oracle (oracle/ref_*.py) and to
generate tests/golden/*.npz (scripts/make_golden.py).  `/root/reference` does
not exist on the GPU box, so nothing at run time may import this module.
"""
import importlib
import os
import sys
import types

REF_MAIN = "/root/reference/main"


def available() -> bool:
    return os.path.isdir(REF_MAIN)


def _stub_config():
    m = types.ModuleType("config")

    class Config:  # values: config.py:3-42
        def __init__(self):
            self.keras_checkpoint_dir = "checkpoints/keras"
            self.trt_checkpoint_dir = "checkpoints/trt"
            self.tensorboard_log_dir = "logs/fit"
            self.trt_precision_mode = "FP32"
            self._N, self._M, self._L = 8, 14, 7
            self.T = 8
            self.num_actions = 4672
            self.conv_filters = 48
            self.num_residual_blocks = 4
            self.input_dims = (8, 8, 14 * self.T + 7)
            self.output_dims = (4672,)
            self.l2_reg = 4e-4
            self.learning_rate = {0: 1e-1, 14000: 1e-2, 28000: 1e-3, 36000: 1e-4}
            self.momentum = 0.9
            self.num_mcts_sims = (50, 300)
            self.num_mcts_sims_p = 0.25
            self.num_mcts_sampling_moves = 30
            self.max_game_length = 512
            self.buffer_size = 1024
            self.minimum_buffer_size = 768
            self.batch_size = 64
            self.training_steps = 40000
            self.pause_between_steps = 2
            self.checkpoint_interval = 2000
            self.epochs = 1
            self.verbose = 1
            self.allow_gpu_growth = True

    m.Config = Config
    return m


def _stub_model():
    m = types.ModuleType("model")

    def predict_fn(trt_func, image):  # model.py:73-80 without tf.function
        import numpy as np
        predictions = trt_func(np.asarray(image, dtype=np.float32)[None])
        return predictions["value_head"][0][0], predictions["policy_head"][0]

    m.predict_fn = predict_fn
    return m


def load():
    """Returns a namespace with the verbatim reference modules."""
    if not available():
        raise RuntimeError("reference tree not present (only exists in the build container)")
    from oracle import pychess_shim
    saved = {k: sys.modules.get(k) for k in ("chess", "config", "model", "actionspace", "gameimage", "game", "mcts")}
    sys.modules["chess"] = pychess_shim
    sys.modules["config"] = _stub_config()
    sys.modules["model"] = _stub_model()
    for k in ("actionspace", "gameimage", "game", "mcts"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_MAIN)
    try:
        ns = types.SimpleNamespace(
            chess=pychess_shim,
            config=sys.modules["config"],
            model=sys.modules["model"],
            actionspace=importlib.import_module("actionspace"),
            gameimage=importlib.import_module("gameimage"),
            game=importlib.import_module("game"),
            mcts=importlib.import_module("mcts"),
        )
    finally:
        sys.path.remove(REF_MAIN)
    ns._saved = saved
    return ns
