"""Self-play at scale on one GPU (BASELINE config 3 shape): `games` concurrent games of the 20x256 net with
the reference's playout-cap randomisation (train.py:29-31), played to the end or to `--max-len` plies.  Reports games/s,
moves/s, simulations/s and replay samples/s, and checks the replay format."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="20x256")
ap.add_argument("--games", type=int, default=512)
ap.add_argument("--max-len", type=int, default=80)
ap.add_argument("--actors", type=int, default=1, help="actor threads sharing the GPU (selfplay.run_actors), games split among them")
a = ap.parse_args()

import bench
from chessbot_b200 import selfplay
from chessbot_b200.config import Config
from chessbot_b200.model import DeviceNetwork, keras_init_weights

blocks, filters = (int(x) for x in a.net.split("x"))
cfg = Config()
cfg.max_game_length = a.max_len
weights = bench.trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks)
rngs = [np.random.default_rng(1000 + g) for g in range(a.games)]
if a.actors == 1:
    net = DeviceNetwork(weights, filters, blocks, max_batch=a.games)
    t0 = time.perf_counter()
    res = selfplay.play_games(net, a.games, cfg, rngs)
else:
    per = (a.games + a.actors - 1) // a.actors
    shares = [rngs[i * per:(i + 1) * per] for i in range(a.actors)]
    nets = selfplay.run_actors([(lambda engine: DeviceNetwork(weights, filters, blocks, max_batch=per, engine=engine)) for _ in shares])
    t0 = time.perf_counter()
    parts = selfplay.run_actors([(lambda engine, i=i: selfplay.play_games(nets[i], len(shares[i]), cfg, shares[i])) for i in range(a.actors)])
    res = [r for part in parts for r in part]
dt = time.perf_counter() - t0
moves = sum(r[3] for r in res)
samples = sum(len(r[0]) for r in res)
outcomes = {}
for r in res:
    outcomes[r[2]] = outcomes.get(r[2], 0) + 1
for imgs, (zs, pis), outcome, hl in res[:8]:
    assert len(imgs) == len(zs) == len(pis) and all(i.shape == (8, 8, 119) and i.dtype == np.int16 for i in imgs)
    assert all(abs(p.sum() - 1) < 1e-9 and p.shape == (4672,) for p in pis) and all(z in (-1, 0, 1) for z in zs)
sims = moves * (0.25 * cfg.num_mcts_sims[1] + 0.75 * cfg.num_mcts_sims[0])
print(f"{a.games} games x {a.net}, {a.actors} actor(s), max {a.max_len} plies: {dt:.1f} s, {moves} moves ({moves / dt:.0f}/s), ~{sims / dt:.0f} sims/s, "
      f"{samples} replay samples ({samples / dt:.0f}/s), outcomes {outcomes}")
