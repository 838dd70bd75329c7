"""Generates tests/golden/*.npz by running the REFERENCE'S OWN FILES (imported verbatim from
/root/reference/main through oracle/verbatim_reference.py) in the build container.

    python scripts/make_golden.py

The fixtures travel to the GPU box; /root/reference does not.  Re-run only here.
Contents:
  actionspace.npz  uci_to_action for all 2*64*63*5 inputs (actionspace.py:116-180), -1 = None
  gameimage.npz    board_to_image (gameimage.py:149-192) for seeded random playouts + the
                   reference's own test FENs; positions stored as FEN + uci move list
  mcts.npz         run_mcts (mcts.py:38-77) visit counts / W / Q / P for seeded searches with the
                   synthetic network (oracle/synthetic_net.py) and pre-drawn Gamma noise
  movegen.json     ordered legal-move golden list of main.ipynb:109-151
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pychess_shim as chess  # noqa: E402
from oracle import synthetic_net, verbatim_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ns = verbatim_reference.load()
cfg = ns.config.Config()


def golden_actionspace():
    table = np.full((2, 64, 64, 5), -1, dtype=np.int16)
    for pi, player in enumerate((True, False)):
        for f in range(64):
            for t in range(64):
                if f == t:
                    continue
                for k, p in enumerate(("", "q", "r", "b", "n")):
                    a = ns.actionspace.uci_to_action(chess.SQUARE_NAMES[f] + chess.SQUARE_NAMES[t] + p, player)
                    if a is not None:
                        table[pi, f, t, k] = int(a)
    np.savez_compressed(os.path.join(OUT, "actionspace.npz"), table=table,
                        moves=np.array(list(ns.actionspace.moves), dtype="U8"))


REF_TEST_FENS = [  # tests/gameimage.py:16-32, tests/mcts_v2.py:12
    "1r1k4/pbpqbp1N/np3n1p/3pp3/Q1N1P3/2PPB3/PP2BPrP/R4R1K w - - 2 15",
    "1r1k1N2/pbpqbp2/np3n1p/3pp3/Q1N1P3/2PPB3/PP2BPrP/R4R1K b - - 3 15",
    "r2qk2r/pppb1ppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPPB1PPP/R2QK2R w KQkq - 4 7",
    "r3k1r1/pppbqppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPPBQPPP/R3K1R1 w Qq - 8 9",
    "1r2k2r/pppbqppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPPBQPPP/1R2K2R w Kk - 8 9",
    "r2k3r/pppbqppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPPBQPPP/R2K3R w - - 8 9",
    "7k/6p1/8/8/8/8/1P6/K7 w - - 1 1",
]


def golden_gameimage():
    rnd = random.Random(20240130)
    fens, movelists, images = [], [], []

    def add(fen, moves):
        b = chess.Board(fen)
        for m in moves:
            b.push_uci(m)
        fens.append(fen)
        movelists.append(" ".join(moves))
        images.append(ns.gameimage.board_to_image(b, 8).astype(np.int16))

    add(chess.STARTING_FEN, [])
    for fen in REF_TEST_FENS:
        add(fen, [])
    # knight shuffles: repetition planes 12/13 in several frames
    add(chess.STARTING_FEN, "g1f3 g8f6 f3g1 f6g8 g1f3 g8f6 f3g1 f6g8 g1f3".split())
    add(chess.STARTING_FEN, "e2e4 e7e5 g1f3 b8c6 f3g1 c6b8 g1f3 b8c6 f3g1 c6b8 g1f3".split())
    for g in range(40):
        b = chess.Board()
        moves = []
        for ply in range(rnd.randint(1, 200)):
            ms = list(b.legal_moves)
            if not ms or b.is_game_over(claim_draw=True):
                break
            # bias towards shuffling so repetition planes fire
            back = [m for m in ms if len(b.move_stack) >= 2 and m.from_square == b.move_stack[-2].to_square
                    and m.to_square == b.move_stack[-2].from_square]
            m = back[0] if back and rnd.random() < 0.35 else rnd.choice(ms)
            b.push(m)
            moves.append(m.uci())
        add(chess.STARTING_FEN, moves)
    imgs = np.stack(images)
    print("gameimage:", imgs.shape, "rep2 frames", int((imgs[:, 0, 0, 12:112:14] > 0).sum()),
          "rep3 frames", int((imgs[:, 0, 0, 13:112:14] > 0).sum()))
    np.savez_compressed(os.path.join(OUT, "gameimage.npz"), fens=np.array(fens), moves=np.array(movelists), images=imgs)


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


MCTS_CASES = [  # (fen, uci prefix, sims, add_noise, CPUCT, seed)
    (chess.STARTING_FEN, "", 100, True, None, 1),
    (chess.STARTING_FEN, "e2e4 e7e5 f1c4 b8c6 d1h5 g8f6", 300, False, 1.0, 2),  # main.ipynb:77-92 position
    ("7k/6p1/8/8/8/8/1P6/K7 w - - 1 1", "", 200, True, None, 3),                 # tests/mcts_v2.py:12
    ("6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1", "", 200, True, None, 4),           # mate in one in the tree
    ("8/8/8/8/8/5k2/6q1/7K b - - 90 100", "", 150, False, None, 5),              # fifty-move claims + mates
    (chess.STARTING_FEN, "g1f3 g8f6 f3g1 f6g8 g1f3 g8f6 f3g1", 120, True, None, 6),  # threefold claim-ahead
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", "", 400, True, None, 7),
    ("8/P6k/8/8/8/8/6p1/K7 w - - 0 1", "", 150, True, None, 8),                  # promotions / underpromotions
]


def golden_mcts():
    out = {}
    for i, (fen, prefix, sims, add_noise, cpuct, seed) in enumerate(MCTS_CASES):
        b = chess.Board(fen)
        for m in prefix.split():
            b.push_uci(m)
        game = ns.game.Game(cfg, board=b)
        noise = np.random.default_rng(seed).gamma(0.3, 1, 256)
        orig = np.random.default_rng
        np.random.default_rng = lambda *a: _FixedRng(noise) if not a else orig(*a)
        try:
            move, root = ns.mcts.run_mcts(game, synthetic_net.as_network(seed), sims, 30, add_noise, cpuct)
        finally:
            np.random.default_rng = orig
        kids = list(root.children.items())
        out[f"c{i}_moves"] = np.array([m for m, _ in kids])
        out[f"c{i}_N"] = np.array([c.N for _, c in kids], dtype=np.int32)
        out[f"c{i}_W"] = np.array([np.float64(c.W) for _, c in kids])
        out[f"c{i}_Q"] = np.array([np.float64(c.Q) for _, c in kids])
        out[f"c{i}_P"] = np.array([np.float64(c.P) for _, c in kids])
        out[f"c{i}_rootN"] = np.int32(root.N)
        out[f"c{i}_move"] = np.array(move)
        # total nodes in the tree, as a checksum of the expansion pattern
        def count(n):
            return 1 + sum(count(c) for c in n.children.values())
        out[f"c{i}_nodes"] = np.int32(count(root))
        print("mcts case", i, move, root.N, out[f"c{i}_nodes"])
    out["cases"] = np.array(json.dumps(MCTS_CASES))
    np.savez_compressed(os.path.join(OUT, "mcts.npz"), **out)


def golden_movegen():
    gold = ("h5h7 h5f7 h5h6 h5g6 h5g5 h5f5 h5e5 h5h4 h5g4 h5h3 h5f3 h5e2 h5d1 c4f7 c4e6 c4a6 c4d5 c4b5 c4d3 c4b3 "
            "c4e2 c4f1 g1h3 g1f3 g1e2 e1e2 e1f1 e1d1 b1c3 b1a3 h2h3 g2g3 f2f3 d2d3 c2c3 b2b3 a2a3 h2h4 g2g4 f2f4 "
            "d2d4 b2b4 a2a4").split()  # main.ipynb:109-151 (root.children order printed by the reference)
    json.dump({"prefix": "e2e4 e7e5 f1c4 b8c6 d1h5 g8f6", "fen": "r1bqkb1r/pppp1ppp/2n2n2/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4",
               "legal_moves_in_order": gold}, open(os.path.join(OUT, "movegen.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_actionspace()
    golden_gameimage()
    golden_mcts()
    golden_movegen()
    print("golden fixtures written to", OUT)
