"""Host logic of the self-play driver that needs no GPU: the visit-distribution gather through the action table equals the
reference's per-child loop (train.py:13-19), promotions and both colours included."""
import numpy as np

from chessbot_b200 import mcts, selfplay
from chessbot_b200.board import Board
from chessbot_b200.config import Config


def test_search_statistics_from_read_back_arrays_equal_the_per_child_loop():
    cfg = Config()
    rng = np.random.default_rng(1)
    for fen, white in (("r1bqkb1r/pppp1ppp/2n2n2/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4", True),
                       ("rnbqkbnr/pppp1ppp/8/4p3/4P3/8/PPPP1PPP/RNBQKBNR b KQkq - 0 2", False),
                       ("8/P7/8/8/8/8/7k/K7 w - - 0 1", True), ("k7/8/8/8/8/8/p7/7K b - - 0 1", False),
                       ("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1", True)):
        b = Board(fen)
        codes = np.array(b.legal_moves._codes(), dtype=np.uint16)
        k = len(codes)
        counts = rng.integers(0, 20, k).astype(np.int32)
        counts[0] += 1
        z = np.zeros(k, np.float32)
        root = mcts._DeviceRoot(int(counts.sum()) + 1, codes, counts, z, z, z)
        fast = selfplay.search_statistics_from_root(cfg, root, white)
        assert "_children" not in root.__dict__                      # the fast path did not build the children
        slow = selfplay.calc_search_statistics(cfg, root, white)
        assert np.array_equal(fast, slow) and abs(fast.sum() - 1) < 1e-12
        # and the lazily built children are the reference's dict: legal-move order, uci keys
        assert list(root.children) == [m.uci() for m in b.legal_moves]


def test_batched_move_choice_equals_the_per_game_reference_calls():
    """mcts._select_moves (one vectorised arg-max for the choices that draw no randomness, the reference's own Generator calls for the
    rest, in game order) against select_move's calls made game by game with an identical generator."""
    import numpy as np
    from chessbot_b200 import mcts
    from chessbot_b200.board import Board
    rng0 = np.random.default_rng(11)
    boards = []
    b = Board()
    for u in ["e2e4", "e7e5", "g1f3", "b8c6", "f1c4", "g8f6", "d2d3", "f8c5"]:
        boards.append(b.copy())
        b.push_uci(u)
    out = {"moves": [], "N": [], "W": [], "Q": [], "P": [], "n_children": [], "root_N": [], "nodes": [], "off": [0]}
    for i, bd in enumerate(boards):
        codes = np.array([m.code() for m in bd.legal_moves], dtype=np.uint16)
        k = len(codes)
        counts = rng0.multinomial(60, rng0.dirichlet(np.full(k, 0.4))).astype(np.int32)
        if i % 3 == 0:
            counts[counts.argmax() - 1] = counts.max()                  # a tie among the most visited
        out["moves"].append(codes); out["N"].append(counts)
        for key, dt in (("W", np.float32), ("Q", np.float32), ("P", np.float64)):
            out[key].append(np.zeros(k, dt))
        out["n_children"].append(k); out["root_N"].append(int(counts.sum()) + 1); out["nodes"].append(0)
        out["off"].append(out["off"][-1] + k)
    packed = {k: (np.concatenate(v) if k in ("moves", "N", "W", "Q", "P") else np.asarray(v, dtype=np.int32)) for k, v in out.items()}
    packed["noisy"] = [True] * len(boards)
    roots = mcts._roots_from_device(packed, len(boards))
    temps = [1.0 if i < 3 else 0 for i in range(len(boards))]
    flat = (packed["moves"], packed["N"], packed["off"], packed["n_children"])
    got = mcts._select_moves(roots, temps, np.random.default_rng(5), flat)
    got_gathered = mcts._select_moves(roots, temps, np.random.default_rng(5), None)
    rng = np.random.default_rng(5)
    want = [mcts._select_move_from_counts(r._stats[0], r._stats[1], rng, t) for r, t in zip(roots, temps)]
    assert got == want and got_gathered == want
    # lazily built children carry the packed statistics in legal-move order
    kids = roots[2].children
    assert list(kids) == [m.uci() for m in boards[2].legal_moves] and [c.N for c in kids.values()] == out["N"][2].tolist()
