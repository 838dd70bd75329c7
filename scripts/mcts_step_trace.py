"""Per-step device time of the four kernels of a search step (CUDA events through cb_profile), step by step."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from chessbot_b200.engine import EVAL_NET, Engine
from chessbot_b200.model import DeviceNetwork, keras_init_weights

net_s = sys.argv[1] if len(sys.argv) > 1 else "20x256"
trees = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 80
filters, blocks = bench.parse_net(net_s)
eng = Engine.get(0)
net = DeviceNetwork(bench.trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks), filters, blocks, max_batch=trees)
boards = bench.random_positions_device(trees, seed=1000)
records = [b.export_record() for b in boards]
eng.search_create(trees, max_nodes=trees * (steps + 6) * 48, max_positions=trees * (steps + 200), max_sims=800)
rng = np.random.default_rng(7)
noise = [rng.gamma(0.3, 1, len(b.legal_moves)) for b in boards]
for rep in range(2):
    eng.search_begin(records, 800, noise=noise, eval_mode=EVAL_NET)
    eng.profile(True)
    rows = []
    for s in range(steps):
        eng.search_run(1)
        p = eng.profile_read()
        rows.append((s, p["mcts_step"][0] * 1e3, p["encode"][0] * 1e3, p["conv"][0] * 1e3, p["heads"][0] * 1e3))
    eng.profile(False)
    st = eng.search_status()
    print(f"rep {rep}: evals {st['evals']} sims {st['sims']}")
    for s, m, e, c, h in rows:
        if s < 12 or s % 8 == 0:
            print(f"  step {s:3d}: mcts {m:7.1f} us  encode {e:5.1f}  tower {c:7.1f}  heads {h:5.1f}")
