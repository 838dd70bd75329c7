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
