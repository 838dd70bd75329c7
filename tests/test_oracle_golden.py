"""The restated oracle (oracle/ref_*.py) against fixtures produced by the reference's OWN files
imported verbatim (scripts/make_golden.py): every action index, 50 plane images, 8 searches."""
import json
import os

import numpy as np
import pytest

from oracle import pychess_shim as chess
from oracle import ref_actionspace, ref_gameimage, ref_mcts, synthetic_net
from oracle.ref_game import Game, OracleConfig


def test_actionspace_exhaustive(golden_dir):
    g = np.load(os.path.join(golden_dir, "actionspace.npz"))
    table = g["table"]
    assert list(g["moves"]) == list(ref_actionspace.moves) and len(ref_actionspace.moves) == 73
    assert len(ref_actionspace.action_space) == 8 * 8 * 73
    n_valid = 0
    for pi, player in enumerate((True, False)):
        for f in range(64):
            for t in range(64):
                if f == t:
                    continue
                for k, p in enumerate(("", "q", "r", "b", "n")):
                    a = ref_actionspace.uci_to_action(chess.SQUARE_NAMES[f] + chess.SQUARE_NAMES[t] + p, player)
                    assert (-1 if a is None else a) == table[pi, f, t, k]
                    n_valid += a is not None
    assert n_valid > 0


def test_actionspace_reference_known_answers():
    # tests/actionspace.py:15-60 (samples of its 200 known answers), both colours
    s = ref_actionspace.uci_to_actionstr
    assert s("a1a2", True) == "Q,N,1" and s("a1b1", True) == "Q,E,1" and s("a1b2", True) == "Q,NE,1"
    assert s("h8h7", True) == "Q,S,1" and s("h8g8", True) == "Q,W,1" and s("h8g7", True) == "Q,SW,1"
    assert s("a8b7", True) == "Q,SE,1" and s("h1g2", True) == "Q,NW,1"
    assert s("d3b2", True) == "K,W,S" and s("d3b4", True) == "K,W,N"
    assert s("a7a8q", True) == "Q,N,1" and s("a7a8r", True) == "P,N,R" and s("a7b8n", True) == "P,NE,K"
    assert s("h2h1b", False) == "P,N,B" and s("h8h7", False) == "Q,N,1"
    assert ref_actionspace.uci_to_action("e1g1", True) == 2346  # SURVEY Appendix A probe
    assert ref_actionspace.uci_to_action("a1b4", True) is None


def _board(fen, moves):
    b = chess.Board(str(fen))
    for m in str(moves).split():
        b.push_uci(m)
    return b


def test_gameimage_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "gameimage.npz"))
    assert len(g["fens"]) == 50
    for fen, moves, img in zip(g["fens"], g["moves"], g["images"]):
        out = ref_gameimage.board_to_image(_board(fen, moves), 8)
        assert out.dtype == np.int16 and out.shape == (8, 8, 119)
        assert np.array_equal(out, img), (fen, moves)


def test_gameimage_shapes_and_errors():
    b = chess.Board()
    for T, c in ((2, 35), (4, 63), (8, 119)):  # tests/gameimage.py:36-47
        assert ref_gameimage.board_to_image(b, T).shape == (8, 8, c)
    with pytest.raises(ValueError):
        ref_gameimage.board_to_image(b, 0)
    img = ref_gameimage.board_to_image(b, 8)
    assert not img[:, :, :98].any() and img[:, :, 98:110].any()  # tests/gameimage.py:49-103


class _FixedRng:
    def __init__(self, noise):
        self.noise = noise

    def gamma(self, a, b, n):
        return self.noise[:n]

    def choice(self, x):
        return x[0]

    def multinomial(self, n, pi):
        o = np.zeros(len(pi), dtype=int)
        o[int(np.argmax(pi))] = 1
        return o


def test_mcts_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "mcts.npz"))
    cases = json.loads(str(g["cases"]))
    for i, (fen, prefix, sims, add_noise, cpuct, seed) in enumerate(cases):
        game = Game(OracleConfig(), board=_board(fen, prefix))
        noise = np.random.default_rng(seed).gamma(0.3, 1, 256)
        move, tree = ref_mcts.run_mcts(game, synthetic_net.as_network(seed), sims, 30, add_noise, cpuct,
                                       rng=_FixedRng(noise), noise=noise)
        kids = list(tree.children(0))
        assert [tree.move[c] for c in kids] == list(g[f"c{i}_moves"])
        assert [tree.N[c] for c in kids] == list(g[f"c{i}_N"]), i
        for name, arr in (("W", tree.W), ("Q", tree.Q), ("P", tree.P)):
            assert np.array_equal(np.array([np.float64(arr[c]) for c in kids]), g[f"c{i}_{name}"]), (i, name)
        assert tree.N[0] == int(g[f"c{i}_rootN"]) and len(tree.parent) == int(g[f"c{i}_nodes"])
        assert move == str(g[f"c{i}_move"])


def test_mcts_reference_known_answers():
    # tests/mcts_v2.py:50-87,152-208 with the mock network restated as f32 arrays
    game = Game(OracleConfig(), board=chess.Board("7k/6p1/8/8/8/8/1P6/K7 w - - 1 1"))
    moves = ["a1a2", "a1b1", "b2b3", "b2b4"]
    logits = np.random.default_rng(0).random(4672).astype(np.float32)
    for m, v in zip(moves, (0.1, 0.2, 0.3, 0.4)):
        logits[ref_actionspace.uci_to_action(m, True)] = v

    def net(_):
        return {"value_head": np.array([[0.1]], np.float32), "policy_head": logits[None]}

    tree = ref_mcts.Tree()
    tree.N[0] = 1
    value = ref_mcts.expand(tree, 0, game, net)
    assert tree.N[0] == 1 and tree.n_children[0] == 4
    s = np.sum(np.exp([0.1, 0.2, 0.3, 0.4]))
    kids = tree.root_children()
    for m, v in zip(moves, (0.1, 0.2, 0.3, 0.4)):
        assert abs(tree.P[kids[m]] - np.exp(v) / s) < 1e-4
    C = np.log((1 + 19652 + 1) / 19652) + 1.25
    for m in moves:
        assert ref_mcts.ucb(tree, 0, kids[m]) == np.float32(np.float32(C) * tree.P[kids[m]])
    leaf = ref_mcts.select_child(tree, 0)
    assert tree.move[leaf] == "b2b4"
    ref_mcts.backup(tree, leaf, -value)
    assert tree.N[leaf] == 1 and tree.W[leaf] == -value and tree.Q[leaf] == -value
    assert tree.N[0] == 2 and tree.W[0] == value


def test_synthetic_net_is_exact_in_bf16():
    v, p = synthetic_net.outputs(ref_gameimage.board_to_image(chess.Board(), 8), seed=3)
    assert -1 <= v < 1 and p.shape == (4672,) and p.min() >= -4 and p.max() < 4
    assert np.array_equal(p * 32, np.round(p * 32))
    assert len(np.unique(p)) > 200


def test_np_sum_model_matches_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [127, 128, 129, 136, 137, 200, 218]:
        a = np.exp(rng.normal(size=n).astype(np.float32) * 3)
        assert ref_mcts.np_sum_f32_model(a) == np.sum(list(a), dtype=np.float32)
