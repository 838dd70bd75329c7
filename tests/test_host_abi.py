"""CPU suite for the drop-in boundary: the C-ABI library loads and exports every symbol of
include/chessbot_b200.h, and its host rules (one source with the device rules, csrc/rules.cuh) agree
with the oracle.  No compute calls on a GPU here."""
import os
import random
import re

import numpy as np
import pytest

from oracle import pychess_shim as ochess
from oracle import ref_actionspace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cb():
    from chessbot_b200 import build
    build.build()
    from chessbot_b200 import _lib, board
    return _lib, board


def test_library_exports_every_declared_symbol(cb):
    _lib, _ = cb
    header = open(os.path.join(ROOT, "include", "chessbot_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)   # names inside comments are not declarations
    declared = sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) > 40
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in include/chessbot_b200.h but not exported"
    assert set(_lib.EXPORTED) <= set(declared)
    assert b"sm_100a" in _lib.lib.cb_version()


def test_no_cpu_fallback(cb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from chessbot_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine()
    assert cb[0].lib.cb_ctx_create(0) is None and b"CUDA" in cb[0].lib.cb_last_error()


PERFT = [
    (ochess.STARTING_FEN, 4, 197281),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", 4, 4085603),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 5, 674624),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 4, 422333),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 4, 2103487),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", 4, 3894594),
]


@pytest.mark.parametrize("fen,depth,nodes", PERFT)
def test_perft_known_answers(cb, fen, depth, nodes):
    assert cb[1].Board(fen).perft(depth) == nodes


def test_rules_agree_with_oracle_on_random_games(cb):
    """move order, outcome(claim_draw), repetition counts, castling rights, action indices on ~10k positions"""
    _, board = cb
    names = {None: None}
    rnd = random.Random(7)
    seen = {}
    npos = 0
    for g in range(45):
        pb, b = ochess.Board(), board.Board()
        bias = 0.5 if g % 3 == 0 else 0.03
        for ply in range(500):
            for claim in (True, False):
                o, o2 = pb.outcome(claim_draw=claim), b.outcome(claim_draw=claim)
                assert (o.termination.name if o else None) == (o2.termination.name if o2 else None), pb.fen()
                assert (o.winner if o else None) == (o2.winner if o2 else None)
            legal = list(pb.legal_moves)
            assert [m.uci() for m in legal] == [m.uci() for m in b.legal_moves], pb.fen()
            assert pb.fen().rsplit(" ", 1)[0] == b.fen().rsplit(" ", 1)[0]
            for k in (2, 3, 5):
                assert bool(pb.is_repetition(k)) == b.is_repetition(k)
            for c in (True, False):
                assert pb.has_kingside_castling_rights(c) == b.has_kingside_castling_rights(c)
                assert pb.has_queenside_castling_rights(c) == b.has_queenside_castling_rights(c)
            assert pb.halfmove_clock == b.halfmove_clock and len(pb.move_stack) == len(b.move_stack) and pb.turn == b.turn
            for m in legal:
                assert cb[0].lib.cb_move_to_action(m.from_square | (m.to_square << 6) | ((m.promotion or 0) << 12), int(pb.turn)) \
                    == ref_actionspace.uci_to_action(m.uci(), pb.turn)
            npos += 1
            o_true = pb.outcome(claim_draw=True)
            if pb.outcome(claim_draw=False) is not None or (o_true is not None and rnd.random() < 0.5):
                seen[o_true.termination.name] = seen.get(o_true.termination.name, 0) + 1
                break
            back = [m for m in legal if len(pb.move_stack) >= 2 and m.from_square == pb.move_stack[-2].to_square
                    and m.to_square == pb.move_stack[-2].from_square]
            m = back[0] if back and rnd.random() < bias else rnd.choice(legal)
            pb.push(m)
            b.push_uci(m.uci())
            if rnd.random() < 0.02 and len(pb.move_stack) > 2:
                assert pb.pop().uci() == b.pop().uci()
    assert npos > 5000 and len(seen) >= 3, seen


def test_reference_golden_move_order(cb, golden_dir):
    import json
    g = json.load(open(os.path.join(golden_dir, "movegen.json")))
    b = cb[1].Board()
    for m in g["prefix"].split():
        b.push_uci(m)
    assert b.fen() == g["fen"]
    assert [m.uci() for m in b.legal_moves] == g["legal_moves_in_order"]


def test_dropin_actionspace_against_reference_table(cb, golden_dir):
    import chessbot_b200.actionspace as asp
    g = np.load(os.path.join(golden_dir, "actionspace.npz"))
    table = g["table"]
    assert list(asp.moves) == list(g["moves"]) and len(asp.moves_dict) == 73 and len(asp.action_space) == 4672
    for pi, player in enumerate((True, False)):
        for f in range(64):
            for t in range(64):
                if f == t:
                    continue
                for k, p in enumerate(("", "q", "r", "b", "n")):
                    a = asp.uci_to_action(ochess.SQUARE_NAMES[f] + ochess.SQUARE_NAMES[t] + p, player)
                    assert (-1 if a is None else a) == table[pi, f, t, k]
    assert asp.uci_to_actionstr("d3b2", True) == "K,W,S" and asp.action_idx_to_str(0) == "Q,N,1"
    with pytest.raises(ValueError):
        asp.uci_to_action("z9a1", True)
    with pytest.raises(ValueError):
        asp.get_action_idx((0, 0), "Q,N,1")   # same outcome as actionspace.py:68


def test_game_wrapper_known_answers(cb):
    # tests/game.py:37-130 through the drop-in Game over the C-ABI board
    from chessbot_b200.config import Config
    from chessbot_b200.game import Game
    _, board = cb

    def play(sans):
        b = board.Board()
        for s in sans:
            b.push_san(s)
        return Game(Config(), board=b)

    g = play(["e4", "e5", "Bc4", "Nc6", "Qh5", "Nf6", "Qxf7"])
    assert g.terminal() and g.terminal_value(True) == 1 and g.terminal_value(False) == -1
    assert g.outcome_str == "White wins: outcome=CHECKMATE"
    g = play(["e4", "e5", "Nc3", "Bc5", "d3", "Qh4", "Nf3", "Qxf2"])
    assert g.terminal() and g.terminal_value(True) == -1 and g.outcome_str == "Black wins: outcome=CHECKMATE"
    g = play(["e4", "e5", "Bc4", "Nc6", "Qh5", "Nf6"])
    assert not g.terminal() and g.terminal_value(True) == 0
    assert g.history == ["e2e4", "e7e5", "f1c4", "b8c6", "d1h5", "g8f6"]
    c = g.clone()
    g.make_move("h5f7")
    assert c.history_len == 6 and g.history_len == 7 and c.to_play() != g.to_play()
    assert len(g.legal_moves()) == 0
    with pytest.raises(ValueError):
        c.make_move("e2e4")
    assert Game(Config(), board=board.Board("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR")).history == []


def test_export_record_covers_frames_and_repetition_window(cb):
    _, board = cb
    b = board.Board()
    for m in "e2e4 e7e5 g1f3 b8c6 f3g1 c6b8 g1f3 b8c6 f3g1 c6b8 g1f3".split():
        b.push_uci(m)
    rec = b.export_record()
    assert rec.shape == (10, 96)                       # back to e7e5 (last irreversible move), >= 8 frames
    ply = rec[:, 74].astype(int) | (rec[:, 75].astype(int) << 8)
    assert list(ply) == list(range(2, 12)) and rec[0, 80] == 1 and rec[-1, 81] == 3   # irrev_in barrier, occur = 3
    b2 = board.Board("r2qk2r/pppb1ppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPPB1PPP/R2QK2R w KQkq - 4 7")
    assert b2.export_record().shape == (1, 96)


def test_weight_checkpoint_roundtrip(tmp_path):
    # model.save_weights / load_weights: the tensor names and Keras layouts of include/chessbot_b200.h (no GPU needed)
    from chessbot_b200.model import keras_init_weights, load_weights, save_weights
    w = keras_init_weights(48, 4, seed=3)
    path = str(tmp_path / "net.npz")
    save_weights(path, w)
    w2, filters, blocks = load_weights(path)
    assert (filters, blocks) == (48, 4) and sorted(w2) == sorted(w)
    assert all(np.array_equal(w[k], w2[k]) and w2[k].dtype == np.float32 for k in w)
    assert w["stem.w"].shape == (3, 3, 119, 48) and w["policy.fc"].shape == (128, 4672) and w["res0.bn1"].shape == (4, 48)


def test_batched_record_export_equals_per_board_export():
    """cb_boards_export_records (roots of a batched search in one call) == cb_board_export_record per board + legal-move counts"""
    import random
    from chessbot_b200.board import Board, export_records
    rnd = random.Random(4)
    boards = []
    for _ in range(40):
        b = Board()
        for _ in range(rnd.randint(0, 90)):
            if b.is_game_over(claim_draw=True):
                break
            b.push(rnd.choice(list(b.legal_moves)))
        boards.append(b)
    boards.append(Board("8/8/8/4k3/8/8/4K3/7R w - - 12 40"))
    recs, n_legal = export_records(boards)
    for b, r, k in zip(boards, recs, n_legal):
        assert np.array_equal(b.export_record(), r) and int(k) == len(b.legal_moves)


def test_search_workspace_size_is_reported_before_allocation(cb):
    """the caller (PyTorch) owns device memory - the library says what a search pool will take (host computation)"""
    _lib, _ = cb
    f = _lib.lib.cb_search_workspace_bytes
    small, big = f(64, 64 * 102 * 48, 64 * 300, 100, 1), f(1024, 1024 * 802 * 48, 1024 * 1000, 800, 1)
    assert 0 < small < big
    assert big > 1024 * 802 * 48 * 32                      # at least the node pool (32-byte nodes)
    assert f(1024, 1024 * 802 * 48, 1024 * 1000, 800, 8) > big


def test_packed_buffer_sizes_match_their_documented_layouts(cb):
    """host-side size helpers of the packed read-backs (include/chessbot_b200.h): the layouts the Python host slices by"""
    _lib, _ = cb
    n, T = 1024, 31337
    # P f64 | N i32 | W f32 | Q f32 per child, four i32 header rows per tree, moves u16 per child
    assert _lib.lib.cb_search_roots_compact_bytes(n, T) == T * (8 + 4 + 4 + 4) + n * 16 + T * 2
    # value f32 | n_moves i32 | offsets i32 [n + 1], then priors f32 | moves u16 | actions i16 per legal move
    assert _lib.lib.cb_leaf_eval_packed_bytes(n, T) == n * 12 + 4 + T * 8
    assert _lib.lib.cb_search_roots_compact_bytes(0, 0) == 0
