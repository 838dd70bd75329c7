"""Drop-in for the reference's model.py: generate_model / predict_fn / predict_model (model.py:30,73,82).

The Keras ResNet becomes `DeviceNetwork`: float32 master weights in Keras layout on the host, the
forward pass on hand-written sm_100a kernels (csrc/conv.cu tcgen05 tower, csrc/heads.cu).  A
DeviceNetwork is also a callable with the reference's `network` contract (model.py:77-79):
network(f32[n,8,8,119]) -> {"value_head": [n,1], "policy_head": [n,4672]}.
"""
import numpy as np
import torch

from .config import Config

config = Config()
NUM_ACTIONS = 4672


def _glorot(gen, shape, fan_in, fan_out):
    lim = float(np.sqrt(6.0 / (fan_in + fan_out)))
    return ((torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim).numpy()


def keras_init_weights(filters, blocks, in_channels=119, seed=None):
    """Keras random-init state of generate_model() (glorot_uniform kernels, BN gamma 1 / beta 0 /
    moving mean 0 / moving variance 1), in Keras tensor layouts (include/chessbot_b200.h: cb_net_set_tensor)."""
    g = torch.Generator()
    if seed is None:
        g.seed()
    else:
        g.manual_seed(seed)
    C = filters

    def bn(c):
        return np.stack([np.ones(c), np.zeros(c), np.zeros(c), np.ones(c)]).astype(np.float32)

    w = {"stem.w": _glorot(g, (3, 3, in_channels, C), 9 * in_channels, 9 * C), "stem.bn": bn(C)}
    for i in range(blocks):
        w[f"res{i}.w1"] = _glorot(g, (3, 3, C, C), 9 * C, 9 * C)
        w[f"res{i}.bn1"] = bn(C)
        w[f"res{i}.w2"] = _glorot(g, (3, 3, C, C), 9 * C, 9 * C)
        w[f"res{i}.bn2"] = bn(C)
    w["value.conv"] = _glorot(g, (1, 1, C, 1), C, 1)
    w["value.bn"] = bn(1)
    w["value.fc1"] = _glorot(g, (64, C), 64, C)
    w["value.fc2"] = _glorot(g, (C, 1), C, 1)
    w["policy.conv"] = _glorot(g, (1, 1, C, 2), C, 2)
    w["policy.bn"] = bn(2)
    w["policy.fc"] = _glorot(g, (128, NUM_ACTIONS), 128, NUM_ACTIONS)
    return w


class DeviceNetwork:
    """model.py:30-56 forward graph resident on one B200."""

    def __init__(self, weights, filters, blocks, max_batch=1024, device=None, engine=None, stream_hi_lo=True):
        """engine: a device context of its own (engine.Engine(device)) instead of the per-GPU default one - one per actor
        thread when several actors share a GPU (the reference runs num_actors processes per GPU, train.py:173,198-201)."""
        from .engine import Engine
        self.weights = {k: np.asarray(v, dtype=np.float32) for k, v in weights.items()}
        self.filters, self.blocks, self.max_batch = filters, blocks, max_batch
        self.stream_hi_lo = stream_hi_lo                # residual stream as bf16 hi + lo between blocks (engine.load_network): bool, or k >= 2
        self.engine = Engine.get(device) if engine is None else engine
        self.load()

    def load(self):
        """(Re)installs these weights as the engine's network (one network per device context)."""
        self.engine.load_network(self.weights, self.filters, self.blocks, self.max_batch, stream_hi_lo=self.stream_hi_lo)
        self.engine.loaded_network = self
        for sib in self.__dict__.get("_siblings", {}).values():     # the copies mcts.SearchStream's other lanes evaluate with follow (fit(), new weights)
            if sib.weights is not self.weights:
                sib.weights = self.weights
                sib.load()

    def sibling(self, index):
        """The same network resident in extra device context number `index` of this GPU (Engine.actor): lets several batched
        searches be in flight at once (mcts.SearchStream), each with its own node pool and activations."""
        sibs = self.__dict__.setdefault("_siblings", {})
        if index not in sibs:
            from .engine import Engine
            sibs[index] = DeviceNetwork(self.weights, self.filters, self.blocks, self.max_batch,
                                        engine=Engine.actor(index, self.engine.device_index), stream_hi_lo=self.stream_hi_lo)
        return sibs[index]

    def ensure_loaded(self):
        if getattr(self.engine, "loaded_network", None) is not self:
            self.load()

    def images_to_planes(self, images):
        """f32 / int16 [n,8,8,119] host images -> tile-native bf16 planes on the device (128 channels; the
        move-count / halfmove-clock parts bf16 cannot hold go to channels 119 / 120, as in csrc/encode.cu)."""
        from .engine import nhwc_to_act
        x = torch.as_tensor(np.asarray(images, dtype=np.float32)).to(self.engine.device)
        n = x.shape[0]
        full = torch.zeros(n, 8, 8, 128, dtype=torch.float32, device=x.device)
        full[..., :119] = x
        for src, dst in ((113, 119), (118, 120)):
            hi = (full[..., src].view(torch.int32) & -65536).view(torch.float32)   # truncate to 8 significant bits
            full[..., dst] = full[..., src] - hi
            full[..., src] = hi
        return nhwc_to_act(full), n

    def predict(self, images):
        """value f32 [n], logits f32 [n,4672] (device tensors) for host images [n,8,8,119]."""
        self.ensure_loaded()
        planes, n = self.images_to_planes(images)
        value, _, logits = self.engine.forward(planes, n, want_bf16=False, want_f32=True)
        return value, logits

    def __call__(self, image):
        value, logits = self.predict(image)
        return {"value_head": value.cpu().numpy().reshape(-1, 1), "policy_head": logits.cpu().numpy()}

    # ------------------------------------------------------------------ the Keras calls train.py makes on its model
    def fit(self, x, y, batch_size=None, epochs=1, initial_epoch=0, shuffle=True, callbacks=None, verbose=0, rng=None):
        """model.fit(x, {"value_head": yv, "policy_head": yp}, initial_epoch=.., epochs=.., callbacks=[lr_callback, ..]) of
        train.py:150-164 with the compiled losses / optimizer of model.py:59-70 (chessbot_b200.train.Trainer underneath; the
        optimizer's momentum persists between calls, like the Keras model's).  batch_size defaults to Keras' 32; one learning
        rate per epoch from a LearningRateScheduler-like callback (anything with .schedule(epoch, lr)), else from
        config.learning_rate (train.py:112-115).  A last partial minibatch is dropped (Keras would run it at its smaller size).
        Afterwards this network evaluates with the trained weights.  Returns an object with .history like Keras."""
        from .train import Trainer, scheduler
        bs = min(32 if batch_size is None else int(batch_size), len(x))
        tr = getattr(self, "_trainer", None)
        if tr is None or tr.batch_size != bs:
            # a trainer is built for one batch size; a new one inherits weights, optimizer momentum, step count and learning rate
            old = tr
            tr = self._trainer = Trainer(self.weights if old is None else old.get_weights(), self.filters, self.blocks, bs,
                                         device=self.engine.device_index)
            if old is None:
                self._lr = tr.config.learning_rate[0]
            else:
                tr.set_weights(old.get_velocity(), what=2)
                tr.steps = old.steps
        history = {"loss": [], "value_head_loss": [], "policy_head_loss": []}
        for epoch in range(initial_epoch, epochs):
            schedules = [cb.schedule for cb in (callbacks or []) if hasattr(cb, "schedule")]
            self._lr = schedules[0](epoch, self._lr) if schedules else scheduler(epoch, self._lr, tr.config)
            logs = tr.fit(x, y, self._lr, batch_size=bs, shuffle=shuffle, rng=rng)
            if logs:
                history["loss"].append(logs["loss"])
                history["value_head_loss"].append(logs["value"])
                history["policy_head_loss"].append(logs["policy"])
                if verbose:
                    print(f"Epoch {epoch + 1}/{epochs} - loss: {logs['loss']:.4f} - value_head_loss: {logs['value']:.4f} - "
                          f"policy_head_loss: {logs['policy']:.4f} - lr: {self._lr:g}")
        self.weights = tr.get_weights()
        self.load()
        return type("History", (), {"history": history, "epoch": list(range(initial_epoch, epochs))})()

    def save(self, path):
        """model.save(path) (train.py:146-148): the .npz of save_weights, loadable with load_network(path)"""
        save_weights(path, self.weights)


def save_weights(path, weights):
    """Checkpoint interchange (train.py:88-101,124-136 write Keras / SavedModel files, unreadable without TensorFlow):
    the float32 master tensors under the names of include/chessbot_b200.h (cb_net_set_tensor) in one .npz.  A Keras
    model exports with  {"stem.w": conv2d.kernel, "stem.bn": [gamma, beta, moving_mean, moving_variance], "res{i}.w1": ...}
    — kernels HWIO and Dense [in, out] exactly as Keras stores them."""
    np.savez(path, **{k: np.asarray(v, dtype=np.float32) for k, v in weights.items()})


def load_weights(path):
    """(weights dict, filters, blocks) from a save_weights() file."""
    with np.load(path) as z:
        w = {k: np.asarray(z[k], dtype=np.float32) for k in z.files}
    filters = int(w["stem.w"].shape[-1])
    blocks = sum(1 for k in w if k.endswith(".w1"))
    return w, filters, blocks


def load_network(path, max_batch=1024, device=None):
    """A DeviceNetwork from a save_weights() .npz, or from the reference's own checkpoints: `model.save("….keras")` archives,
    `.weights.h5` / legacy `.h5` files (train.py:124-136, init.py:44) through chessbot_b200.keras_io (no TensorFlow needed)."""
    if str(path).endswith((".keras", ".h5", ".hdf5")):
        from .keras_io import load_keras
        w, filters, blocks = load_keras(path)
        w["value.conv"], w["policy.conv"] = w["value.conv"].reshape(-1, 1), w["policy.conv"].reshape(-1, 2)
        return DeviceNetwork(w, filters, blocks, max_batch, device)
    w, filters, blocks = load_weights(path)
    return DeviceNetwork(w, filters, blocks, max_batch, device)


class SyntheticNetwork:
    """The reference tests' mock network (tests/mcts_v2.py:51-67) as a device kernel: outputs are an
    integer hash of the planes (csrc/heads.cu; NumPy statement in oracle/synthetic_net.py)."""

    def __init__(self, seed=0, device=None):
        from .engine import Engine
        self.seed = int(seed)
        self.engine = Engine.get(device)

    def __call__(self, image):
        planes = torch.as_tensor(np.asarray(image).astype(np.int16)).to(self.engine.device).contiguous()
        value, logits = self.engine.synth_forward(planes, self.seed)
        return {"value_head": value.cpu().numpy().reshape(-1, 1), "policy_head": logits.float().cpu().numpy()}


def generate_model(seed=None, max_batch=1024):
    """model.py:30-71: a randomly initialised network of config.conv_filters x config.num_residual_blocks."""
    w = keras_init_weights(config.conv_filters, config.num_residual_blocks, config.input_dims[2], seed)
    return DeviceNetwork(w, config.conv_filters, config.num_residual_blocks, max_batch)


def predict_fn(trt_func, image):
    """model.py:73-80: (value, policy_logits[4672]) for ONE image [8,8,119]; `trt_func` is any callable with
    the network contract (a DeviceNetwork, a SyntheticNetwork or a user mock)."""
    image = np.asarray(image, dtype=np.float32)[None]
    predictions = trt_func(image)
    policy_logits = predictions["policy_head"][0]
    value = predictions["value_head"][0][0]
    return value, policy_logits


def predict_model(model, image):
    """model.py:82-89 (Keras model returning [value, policy])."""
    if isinstance(model, (DeviceNetwork, SyntheticNetwork)):
        return predict_fn(model, image)
    image = np.asarray(image, dtype=np.float32)[None]
    predictions = model(image)
    return predictions[0][0], predictions[1][0]
