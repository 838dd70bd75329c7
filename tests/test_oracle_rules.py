"""Pins the oracle's python-chess restatement (oracle/pychess_shim.py): public perft counts, the
reference's ordered golden move list (main.ipynb:109-151), the known answers of the reference's
tests/game.py (:37-118), and the draw-claim rules restated from python-chess."""
import json
import os

import pytest

from oracle import pychess_shim as chess
from oracle.ref_game import Game, OracleConfig

PERFT = [
    (chess.STARTING_FEN, [20, 400, 8902]),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", [48, 2039, 97862]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", [14, 191, 2812, 43238]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", [6, 264, 9467]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890]),
]


@pytest.mark.parametrize("fen,counts", PERFT)
def test_perft(fen, counts):
    b = chess.Board(fen)
    assert [chess.perft(b, d + 1) for d in range(len(counts))] == counts


def test_reference_golden_move_order(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "movegen.json")))
    b = chess.Board()
    for m in g["prefix"].split():
        b.push_uci(m)
    assert b.fen() == g["fen"]
    assert [m.uci() for m in b.legal_moves] == g["legal_moves_in_order"]


def _play(sans):
    b = chess.Board()
    for s in sans:
        b.push_san(s)
    return Game(OracleConfig(), board=b)


def test_reference_game_known_answers():
    # tests/game.py:37-118
    g = _play(["e4", "e5", "Bc4", "Nc6", "Qh5", "Nf6", "Qxf7"])
    assert g.terminal() and g.terminal_value(chess.WHITE) == 1 and g.terminal_value(chess.BLACK) == -1
    assert g.outcome_str == "White wins: outcome=CHECKMATE"
    g = _play(["e4", "e5", "Nc3", "Bc5", "d3", "Qh4", "Nf3", "Qxf2"])
    assert g.terminal() and g.terminal_value(chess.WHITE) == -1 and g.terminal_value(chess.BLACK) == 1
    g = _play(["e4", "e5", "Bc4", "Nc6", "Qh5", "Nf6"])
    assert not g.terminal() and g.terminal_value(chess.WHITE) == 0
    assert g.history == ["e2e4", "e7e5", "f1c4", "b8c6", "d1h5", "g8f6"] and g.history_len == 6
    c = g.clone()
    g.make_move("h5f7")
    assert c.history_len == 6 and g.history_len == 7 and c.to_play() != g.to_play()
    assert Game(OracleConfig(), board=chess.Board("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR")).history == []


def test_four_legal_moves_fixture():
    # tests/mcts_v2.py:12,78
    b = chess.Board("7k/6p1/8/8/8/8/1P6/K7 w - - 1 1")
    assert [m.uci() for m in b.legal_moves] == ["a1a2", "a1b1", "b2b3", "b2b4"]
    b.push_uci("b2b4")
    assert [m.uci() for m in b.legal_moves] == ["h8g8", "h8h7", "g7g6", "g7g5"]


def test_threefold_claim_ahead_and_repetition_planes():
    b = chess.Board()
    seq = "g1f3 g8f6 f3g1 f6g8 g1f3 g8f6 f3g1".split()
    for i, m in enumerate(seq):
        assert not b.is_game_over(claim_draw=True), i
        b.push_uci(m)
    # black to move: f6g8 would produce the start position for the third time -> claimable now
    assert b.can_claim_threefold_repetition()
    assert b.outcome(claim_draw=True).termination.name == "THREEFOLD_REPETITION"
    assert not b.is_game_over(claim_draw=False)
    assert b.is_repetition(2) and not b.is_repetition(3)
    b.push_uci("f6g8")
    assert b.is_repetition(3)


def test_fifty_move_claims():
    b = chess.Board("8/8/8/8/8/5k2/6q1/7K b - - 98 100")
    assert not b.is_game_over(claim_draw=True)
    b = chess.Board("8/8/8/3k4/8/8/6q1/K7 b - - 99 100")
    assert b.outcome(claim_draw=True).termination.name == "FIFTY_MOVES"  # claim-ahead
    b = chess.Board("8/8/8/3k4/8/8/6q1/K7 b - - 100 100")
    assert b.outcome(claim_draw=True).termination.name == "FIFTY_MOVES"
    b = chess.Board("8/8/8/3k4/8/8/6q1/K7 b - - 150 100")
    assert b.outcome(claim_draw=False).termination.name == "SEVENTYFIVE_MOVES"


def test_terminations():
    assert chess.Board("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1").outcome().termination.name == "STALEMATE"
    assert chess.Board("7k/8/6K1/8/8/8/8/8 b - - 0 1").outcome().termination.name == "INSUFFICIENT_MATERIAL"
    assert chess.Board("7k/8/6K1/8/8/8/8/B7 b - - 0 1").outcome().termination.name == "INSUFFICIENT_MATERIAL"
    assert chess.Board("7k/8/6K1/8/8/8/8/BB6 b - - 0 1").outcome() is None  # opposite-coloured bishops
    assert chess.Board("7k/8/6K1/8/8/8/8/B1B5 b - - 0 1").outcome().termination.name == "INSUFFICIENT_MATERIAL"
    assert chess.Board("7k/8/6K1/8/8/8/8/NN6 b - - 0 1").outcome() is None
    o = chess.Board("R6k/6pp/8/8/8/8/8/K7 b - - 0 1").outcome()
    assert o.termination.name == "CHECKMATE" and o.winner is True


def test_castling_rights_cleaning_and_en_passant_key():
    b = chess.Board("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1")
    assert [m.uci() for m in b.legal_moves if m.from_square == chess.E1][-2:] == ["e1g1", "e1c1"]
    b.push_uci("h1h2")
    assert not b.has_kingside_castling_rights(True) and b.has_queenside_castling_rights(True)
    b = chess.Board("4k3/8/8/8/8/8/8/R3K2R w KQkq - 0 1")  # black rights cleaned away
    assert b.fen().split()[2] == "KQ"
    # ep square only counts when an en-passant capture is legal
    b = chess.Board()
    b.push_uci("e2e4")
    assert b.ep_square == chess.parse_square("e3") and not b.has_legal_en_passant() and b.fen().split()[3] == "-"
    for m in "a7a6 e4e5 d7d5".split():
        b.push_uci(m)
    assert b.has_legal_en_passant() and "e5d6" in [m.uci() for m in b.legal_moves]
    assert [m.uci() for m in b.legal_moves][-1] == "e5d6"  # en passant is generated last


def test_mirror():
    b = chess.Board()
    b.push_uci("e2e4")
    m = b.mirror()
    assert m.turn == chess.WHITE and m.fen().split()[0] == "rnbqkbnr/pppp1ppp/8/4p3/8/8/PPPPPPPP/RNBQKBNR"
