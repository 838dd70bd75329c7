"""N>1 host logic on CPU: world_size-2 gloo processes shard games by rank (g -> g mod world) and gather
the per-game results in order; the per-game work here is the oracle search (the checker), since the
device search needs a GPU.  Covers chessbot_b200.sharding, which bench.py --gpus N and run_mcts_sharded use."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_search(move_lists):
    from oracle import pychess_shim as chess
    from oracle import ref_mcts, synthetic_net
    from oracle.ref_game import Game, OracleConfig
    out = []
    for moves in move_lists:
        b = chess.Board()
        for m in moves:
            b.push_uci(m)
        mv, tree = ref_mcts.run_mcts(Game(OracleConfig(), board=b), synthetic_net.as_network(1), 6, 0, False, None,
                                     rng=np.random.default_rng(0))
        out.append((mv, [tree.N[c] for c in tree.children(0)]))
    return out


def _worker(rank, world, port, lines, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chessbot_b200 import sharding
    assert sharding.world() == (rank, world)
    full = sharding.sharded_map(_oracle_search, lines)
    idx, local = sharding.sharded_map(_oracle_search, lines, gather=False)
    q.put((rank, full, idx, local))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    from chessbot_b200.sharding import shard_indices
    for n in (0, 1, 7, 512):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert all(i % w == r for r, p in enumerate(parts) for i in p)
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_world_size_2_gloo_sharded_search_matches_single_process():
    lines = [[], ["e2e4"], ["e2e4", "e7e5"], ["d2d4", "d7d5", "c2c4"], ["g1f3"]]
    expected = _oracle_search(lines)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lines, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, idx, local in got:
        assert full == expected                       # every rank holds all results, in game order
        assert idx == list(range(rank, len(lines), 2)) and local == [expected[i] for i in idx]
