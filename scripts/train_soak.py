"""Training soak on one GPU: many SGD steps at the reference's learning rate on a fixed pool of synthetic samples; the losses must
stay finite and fall, the BN running statistics must stay sane, and the trained weights must serve the inference path."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--net", default="4x48")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--pool", type=int, default=2048)
ap.add_argument("--lr", type=float, default=0.1)
a = ap.parse_args()

import bench
from chessbot_b200.gameimage import boards_to_images
from chessbot_b200.model import DeviceNetwork, keras_init_weights
from chessbot_b200.train import Trainer

blocks, filters = (int(x) for x in a.net.split("x"))
boards = bench.random_positions_device(min(a.pool, 512), seed=7)
imgs = boards_to_images(boards, 8).astype(np.int16)
imgs = np.concatenate([imgs] * (a.pool // len(imgs) + 1))[:a.pool]
rng0 = np.random.default_rng(3)
pi = np.zeros((a.pool, 4672), dtype=np.float32)            # sparse visit distributions (train.py:13-19 shapes)
cols = np.stack([rng0.choice(4672, 30, replace=False) for _ in range(a.pool)])
np.put_along_axis(pi, cols, rng0.integers(1, 50, (a.pool, 30)).astype(np.float32), axis=1)
pi /= pi.sum(1, keepdims=True)
z = np.sign(imgs[:, 0, 0, 112] - 0.5).astype(np.float32)                 # a learnable value target: the side to move
tr = Trainer(keras_init_weights(filters, blocks, seed=0), filters, blocks, a.batch)
rng = np.random.default_rng(0)
t0 = time.perf_counter()
hist = []
for s in range(a.steps):
    idx = rng.choice(a.pool, a.batch, replace=False)
    hist.append(tr.train_on_batch(imgs[idx], z[idx], pi[idx], a.lr))
    if not all(np.isfinite(v) for v in hist[-1].values()):
        raise SystemExit(f"step {s}: non-finite loss {hist[-1]}")
dt = time.perf_counter() - t0
for s in range(300):       # weights frozen (lr 0, momentum 0): the BN running statistics settle on the final weights' batch statistics
    idx = rng.choice(a.pool, a.batch, replace=False)
    tr.forward_backward(imgs[idx], z[idx], pi[idx])
    tr.apply(0.0, momentum=0.0)
frozen = tr.losses()
first = {k: float(np.mean([h[k] for h in hist[:10]])) for k in hist[0]}
last = {k: float(np.mean([h[k] for h in hist[-10:]])) for k in hist[0]}
w = tr.get_weights()
var_ok = all(np.all(v[3] > 0) and np.all(np.isfinite(v)) for k, v in w.items() if "bn" in k)
net = DeviceNetwork(w, filters, blocks, max_batch=64)
value, logits = net.predict(imgs[:64])
acc = float(np.mean(np.sign(value.cpu().numpy()) == z[:64]))
print(f"{a.net} batch {a.batch}: {a.steps} steps in {dt:.1f} s ({a.steps * a.batch / dt:.0f} samples/s through train_on_batch from NumPy), "
      f"loss {first['loss']:.3f} -> {last['loss']:.3f} (value {first['value']:.3f} -> {last['value']:.3f}, policy {first['policy']:.3f} -> "
      f"{last['policy']:.3f}; frozen weights: value {frozen['value']:.3f}), BN statistics sane: {var_ok}, inference with the trained weights: "
      f"value sign accuracy {acc:.2f}, inference value mse {float(np.mean((value.cpu().numpy() - z[:64]) ** 2)):.3f}")
assert var_ok and last["loss"] < first["loss"] and last["value"] < 0.5 * first["value"] and acc > 0.95
