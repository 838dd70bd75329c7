"""Small end-to-end invocations of every kernel family, meant for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`
where the tool is available (it is closed on the build pool: the script then just runs, with CB_DEBUG_SYNC=1 naming a faulting
stage): a search with the device network and a shrinking batch, the stand-alone leaf evaluation, the wgrad hook, and training
steps (odd batch, padded channels; eager, graph capture, replay)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from chessbot_b200.board import Board
from chessbot_b200.engine import EVAL_NET, Engine, nhwc_to_act
from chessbot_b200.model import keras_init_weights
from chessbot_b200.train import Trainer

eng = Engine.get(0)
w = keras_init_weights(48, 1, seed=1)
eng.load_network(w, 48, 1, max_batch=8)
boards = []
for line in ("e2e4 e7e5 g1f3", "d2d4", "", "g1f3 g8f6 f3g1 f6g8", "c2c4 e7e5 b1c3"):
    b = Board()
    for m in line.split():
        b.push_uci(m)
    boards.append(b)
recs = [b.export_record() for b in boards]
eng.search_create(max_trees=8, max_nodes=100_000, max_positions=10_000, max_sims=32)
out = eng.search(recs, [12, 5, 9, 12, 3], eval_mode=EVAL_NET)          # trees finish at different steps: the batch shrinks
assert out["sims"] == 41, out["sims"]
eng.evaluate_positions(recs)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(5, 8, 8, 64, device="cuda", generator=g)
dy = torch.randn(5, 8, 8, 64, device="cuda", generator=g)
eng.debug_wgrad(nhwc_to_act(x), nhwc_to_act(dy), 5, 64, 64)
x = torch.randn(3, 8, 8, 256, device="cuda", generator=g)
dy = torch.randn(3, 8, 8, 256, device="cuda", generator=g)
eng.debug_wgrad(nhwc_to_act(x), nhwc_to_act(dy), 3, 256, 256)
rng = np.random.default_rng(0)
for filters, n in ((48, 5), (128, 4)):
    tr = Trainer(keras_init_weights(filters, 1, seed=2), filters, 1, n)
    img = rng.integers(0, 2, (n, 8, 8, 119)).astype(np.int16)
    z = rng.integers(-1, 2, n).astype(np.float32)
    pi = rng.random((n, 4672)).astype(np.float32)
    pi /= pi.sum(1, keepdims=True)
    for _ in range(3):                                                     # eager, capture, replay
        loss = tr.train_on_batch(img, z, pi, 0.05)
    assert np.isfinite(loss["loss"])
torch.cuda.synchronize()
print("sanitize_small ok")
