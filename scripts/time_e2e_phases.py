"""Wall-clock phases of one mcts.run_mcts_batch call (1024 games, 100 sims, 20x256) on the GPU box."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from chessbot_b200 import mcts as cb_mcts
from chessbot_b200.board import export_records
from chessbot_b200.config import Config
from chessbot_b200.engine import EVAL_NET
from chessbot_b200.game import Game
from chessbot_b200.gameimage import as_device_board
from chessbot_b200.model import DeviceNetwork, keras_init_weights

filters, blocks, trees, sims = 256, 20, 1024, 100
net = DeviceNetwork(bench.trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks), filters, blocks, max_batch=trees)
eng = net.engine
games = [Game(Config(), board=b) for b in bench.random_positions_device(trees, seed=1000)]
rng = np.random.default_rng(3)
cb_mcts.run_mcts_batch(games, net, sims, 30, True, None, rng=rng)
for rep in range(2):
    torch.cuda.synchronize()
    T = [time.perf_counter()]
    mark = lambda: T.append(time.perf_counter())
    boards = [as_device_board(g.board) for g in games]
    records, n_legal = export_records(boards)
    flat = rng.gamma(0.3, 1, int(n_legal.sum()))
    ends = np.cumsum(n_legal)
    noise = [flat[e - k:e] for e, k in zip(ends.tolist(), n_legal.tolist())]
    mark()
    eng.search_begin(records, [sims] * trees, cpuct=None, noise=noise, max_game_length=512, eval_mode=EVAL_NET)
    mark()
    eng.search_run(sims + 1)
    mark()
    st = eng.search_status()
    mark()
    out = eng.search_read_roots()
    mark()
    out["noisy"] = [True] * trees
    roots = cb_mcts._roots_from_device(out, trees)
    mark()
    moves = [cb_mcts._select_move_from_counts(r._stats[0], r._stats[1], rng, 1.0 if g.history_len < 30 else 0) for g, r in zip(games, roots)]
    mark()
    for r in roots[::4]:
        sum(c.N for c in r.children.values())
    mark()
    names = ["prep(export+noise)", "search_begin", "enqueue 101 steps", "wait (status sync)", "read_roots", "lazy roots", "select moves", "children of 1/4"]
    print({n: round(1e3 * (b - a), 2) for n, a, b in zip(names, T, T[1:])}, "total", round(1e3 * (T[-1] - T[0]), 1), "evals", st["evals"])
