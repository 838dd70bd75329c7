"""cProfile of the host side of mcts.run_mcts_batch (1024 games, few simulations: device time small) on the GPU box."""
import cProfile
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from chessbot_b200 import mcts as cb_mcts
from chessbot_b200.config import Config
from chessbot_b200.game import Game
from chessbot_b200.model import DeviceNetwork, keras_init_weights

filters, blocks, trees, sims = 48, 4, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 10
net = DeviceNetwork(bench.trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks), filters, blocks, max_batch=trees)
boards = bench.random_positions_device(trees, seed=1000)
games = [Game(Config(), board=b) for b in boards]
cb_mcts.run_mcts_batch(games, net, sims, 30, True, None, rng=np.random.default_rng(3))
t0 = time.perf_counter()
cb_mcts.run_mcts_batch(games, net, sims, 30, True, None, rng=np.random.default_rng(3))
print("one call:", time.perf_counter() - t0, "s")
pr = cProfile.Profile()
pr.enable()
cb_mcts.run_mcts_batch(games, net, sims, 30, True, None, rng=np.random.default_rng(3))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
