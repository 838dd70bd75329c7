"""K5-K8 parity (-m gpu): legal-move gather/softmax and the batched device search through the C ABI.

Bars (BASELINE.json north_star): legal-move indices and visit counts bit-exact given identical network
outputs and seeded noise.  Float semantics = the reference under NumPy 2 with a float32 network
(oracle/ref_mcts.py header): f32 priors / W / Q / ucb, f64 at a noisy root."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from chessbot_b200.engine import Engine
    e = Engine.get()
    e.search_create(max_trees=64, max_nodes=4_000_000, max_positions=400_000, max_sims=800)
    return e


def test_policy_softmax_matches_oracle(engine):
    from gpu_util import device_board, oracle_board, random_games
    from chessbot_b200.board import code_to_uci
    from oracle import ref_actionspace, ref_gameimage, synthetic_net
    games = random_games(96, seed=5, max_plies=100)
    boards = [device_board(f, m) for f, m in games]
    pos, fi = engine.upload_records([b.export_record() for b in boards])
    _, i16 = engine.encode(pos, fi, want_act=False, want_i16=True)
    value, lb = engine.synth_forward(i16, seed=77)
    moves, actions, priors, counts = engine.policy_softmax(pos, fi[:, 0].contiguous(), lb)
    torch.cuda.synchronize()
    moves, actions, priors, counts = (t.cpu().numpy() for t in (moves, actions, priors, counts))
    for i, (f, m) in enumerate(games):
        ob = oracle_board(f, m)
        v_ref, logits_ref = synthetic_net.outputs(ref_gameimage.board_to_image(ob, 8), seed=77)
        assert np.float32(value[i].item()) == v_ref
        assert np.array_equal(lb[i].float().cpu().numpy(), logits_ref)
        legal = [mv.uci() for mv in ob.legal_moves]
        k = counts[i]
        assert [code_to_uci(c) for c in moves[i, :k].astype(np.uint16)] == legal          # order too
        acts = [ref_actionspace.uci_to_action(u, ob.turn) for u in legal]
        assert list(actions[i, :k]) == acts
        p = np.exp(logits_ref[acts], dtype=np.float32)
        ref = p / np.sum(p, dtype=np.float32)
        assert np.array_equal(priors[i, :k], ref)


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


def _compare_tree(out, i, tree, root_moves=None):
    from chessbot_b200.board import code_to_uci
    kids = list(tree.children(0))
    k = out["n_children"][i]
    assert k == len(kids)
    assert [code_to_uci(c) for c in out["moves"][i, :k]] == [tree.move[c] for c in kids]
    assert list(out["N"][i, :k]) == [tree.N[c] for c in kids]
    assert np.array_equal(out["W"][i, :k], np.array([tree.W[c] for c in kids], dtype=np.float32))
    assert np.array_equal(out["Q"][i, :k], np.array([tree.Q[c] for c in kids], dtype=np.float32))
    assert np.array_equal(out["P"][i, :k], np.array([np.float64(tree.P[c]) for c in kids]))
    assert out["root_N"][i] == tree.N[0] and out["nodes"][i] == len(tree.parent)


def test_reference_golden_searches_bit_exact(engine, golden_dir):
    """The eight searches recorded from the reference's own mcts.py (tests/golden/mcts.npz), all run
    CONCURRENTLY as one device batch with per-tree sims / cpuct / noise."""
    from gpu_util import device_board
    from chessbot_b200.board import code_to_uci
    from chessbot_b200.engine import EVAL_SYNTH
    g = np.load(os.path.join(golden_dir, "mcts.npz"))
    cases = json.loads(str(g["cases"]))
    seeds = {c[5] for c in cases}
    for seed in seeds:   # the synthetic network seed is per launch
        idx = [i for i, c in enumerate(cases) if c[5] == seed]
        recs = [device_board(cases[i][0], cases[i][1]).export_record() for i in idx]
        noise = [np.random.default_rng(cases[i][5]).gamma(0.3, 1, 256) if cases[i][3] else None for i in idx]
        out = engine.search(recs, [cases[i][2] for i in idx], cpuct=[cases[i][4] for i in idx], noise=noise,
                            eval_mode=EVAL_SYNTH, seed=seed)
        for j, i in enumerate(idx):
            k = out["n_children"][j]
            assert [code_to_uci(c) for c in out["moves"][j, :k]] == list(g[f"c{i}_moves"])
            assert list(out["N"][j, :k]) == list(g[f"c{i}_N"]), i
            assert np.array_equal(out["W"][j, :k].astype(np.float64), g[f"c{i}_W"]), i
            assert np.array_equal(out["Q"][j, :k].astype(np.float64), g[f"c{i}_Q"]), i
            assert np.array_equal(out["P"][j, :k], g[f"c{i}_P"]), i
            assert out["root_N"][j] == int(g[f"c{i}_rootN"]) and out["nodes"][j] == int(g[f"c{i}_nodes"])


def test_batched_synthetic_search_matches_oracle(engine):
    """48 concurrent trees on random mid-game positions (repetitions, checks, promotions in the trees)."""
    from gpu_util import device_board, oracle_board, random_games
    from chessbot_b200.engine import EVAL_SYNTH
    from oracle import ref_mcts, synthetic_net
    from oracle.ref_game import Game, OracleConfig
    games = random_games(48, seed=21, max_plies=140, back_bias=0.35)
    recs = [device_board(f, m).export_record() for f, m in games]
    sims = [40 + 7 * (i % 9) for i in range(48)]
    noise = [np.random.default_rng(100 + i).gamma(0.3, 1, 256) if i % 2 == 0 else None for i in range(48)]
    cpuct = [None if i % 3 else 1.0 for i in range(48)]
    out = engine.search(recs, sims, cpuct=cpuct, noise=noise, eval_mode=EVAL_SYNTH, seed=9)
    for i, (f, m) in enumerate(games):
        game = Game(OracleConfig(), board=oracle_board(f, m))
        _, tree = ref_mcts.run_mcts(game, synthetic_net.as_network(9), sims[i], 30, noise[i] is not None, cpuct[i],
                                    rng=_FixedRng(noise[i]), noise=noise[i])
        _compare_tree(out, i, tree)
    assert out["sims"] == sum(sims)


def test_device_network_search_matches_oracle_on_same_outputs(engine):
    """Search with the tcgen05 network in the loop.  The oracle gets the SAME network outputs (the
    device forward pass called per position through the C ABI), so visit counts must be identical."""
    from gpu_util import device_board, oracle_board, random_games
    from chessbot_b200.board import Board
    from chessbot_b200.engine import EVAL_NET
    from oracle import ref_mcts, ref_model
    from oracle.ref_game import Game, OracleConfig
    w = ref_model.init_weights(64, 2, seed=12)
    engine.load_network(w, 64, 2, max_batch=64)
    games = random_games(6, seed=33, max_plies=60)
    recs = [device_board(f, m).export_record() for f, m in games]
    sims = 60
    noise = [np.random.default_rng(i).gamma(0.3, 1, 256) for i in range(6)]
    out = engine.search(recs, sims, noise=noise, eval_mode=EVAL_NET)
    for i, (f, m) in enumerate(games):
        ob = oracle_board(f, m)
        game = Game(OracleConfig(), board=ob)

        class DeviceNet:
            """network(image) evaluated by the device forward on the oracle's own game record"""
            def __init__(self, g):
                self.g = g

        def run_with_device_outputs():
            # the oracle asks for image -> outputs; feed it the device's outputs for the same position
            cache = {}

            def expand_hook(tree, node, g, network, exp_fn=ref_mcts.default_exp):
                db = Board(str(f))
                for u in [mv.uci() for mv in g.board.move_stack]:
                    db.push_uci(u)
                pos, fi = engine.upload_records([db.export_record()])
                act, _ = engine.encode(pos, fi)
                v, lb, _ = engine.forward(act, 1)
                torch.cuda.synchronize()
                value = np.float32(v.cpu().numpy()[0])
                logits = lb.float().cpu().numpy()[0]
                moves, actions = ref_mcts.legal_actions(g)
                p = exp_fn(logits[np.asarray(actions, dtype=np.int64)])
                s = np.sum(p, dtype=np.float32)
                tree.add_children(node, moves, [np.float32(x / s) for x in p])
                return value

            orig = ref_mcts.expand
            ref_mcts.expand = expand_hook
            try:
                return ref_mcts.run_mcts(game, None, sims, 30, True, None, rng=_FixedRng(noise[i]), noise=noise[i])
            finally:
                ref_mcts.expand = orig

        _, tree = run_with_device_outputs()
        _compare_tree(out, i, tree)


def test_config1_single_game_start_position_100_sims_default_net(engine):
    """BASELINE config 1: single game from the start position, 100 sims, the repo-default 4x48 net, batch 1 —
    the device tree against the oracle fed with the device network's own outputs (visit counts identical)."""
    from chessbot_b200.board import Board
    from chessbot_b200.engine import EVAL_NET
    from oracle import pychess_shim as ochess
    from oracle import ref_mcts, ref_model
    from oracle.ref_game import Game, OracleConfig
    w = ref_model.init_weights(48, 4, seed=1)
    engine.load_network(w, 48, 4, max_batch=64)
    noise = np.random.default_rng(5).gamma(0.3, 1, 256)
    out = engine.search([Board().export_record()], 100, noise=[noise], eval_mode=EVAL_NET)

    def expand_with_device_outputs(tree, node, g, network, exp_fn=ref_mcts.default_exp):
        db = Board()
        for mv in g.board.move_stack:
            db.push_uci(mv.uci())
        pos, fi = engine.upload_records([db.export_record()])
        act, _ = engine.encode(pos, fi)
        v, lb, _ = engine.forward(act, 1)
        torch.cuda.synchronize()
        logits = lb.float().cpu().numpy()[0]
        moves, actions = ref_mcts.legal_actions(g)
        p = exp_fn(logits[np.asarray(actions, dtype=np.int64)])
        s = np.sum(p, dtype=np.float32)
        tree.add_children(node, moves, [np.float32(x / s) for x in p])
        return np.float32(v.cpu().numpy()[0])

    orig = ref_mcts.expand
    ref_mcts.expand = expand_with_device_outputs
    try:
        _, tree = ref_mcts.run_mcts(Game(OracleConfig(), board=ochess.Board()), None, 100, 30, True, None, rng=_FixedRng(noise), noise=noise)
    finally:
        ref_mcts.expand = orig
    _compare_tree(out, 0, tree)
    assert out["root_N"][0] == 101 and out["sims"] == 100


def test_config3_512_games_800_sims_20x256_properties():
    """BASELINE config 3 at full size (512 concurrent games x 800 simulations, 20x256 net) through size-independent
    properties: every tree finishes with exactly 800 visits below the root, Q == W / N bit for bit, priors sum to
    one, and a tree's statistics do not depend on the batch it was searched in (re-searched alone: identical)."""
    import random
    from chessbot_b200.board import Board
    from chessbot_b200.engine import EVAL_NET, Engine
    from chessbot_b200.model import keras_init_weights
    import bench
    eng = Engine.get()
    eng.load_network(bench.trained_like_bn(keras_init_weights(256, 20, seed=0), 256, 20), 256, 20, max_batch=512)
    eng.search_create(max_trees=512, max_nodes=512 * 802 * 40, max_positions=512 * 1100, max_sims=800)
    rnd = random.Random(3)
    recs = []
    for _ in range(512):                                  # self-play openings: 0..11 random plies from the start position
        b = Board()
        for _ in range(rnd.randint(0, 11)):
            b.push(rnd.choice(list(b.legal_moves)))
        recs.append(b.export_record())
    noise = [np.random.default_rng(i).gamma(0.3, 1, 256) for i in range(512)]
    out = eng.search(recs, 800, noise=noise, eval_mode=EVAL_NET)
    assert out["unfinished"] == 0 and out["errors"] == 0 and out["sims"] == 512 * 800 and out["evals"] <= 512 * 801
    for i in range(512):
        k = out["n_children"][i]
        n, w, q, p = out["N"][i, :k], out["W"][i, :k], out["Q"][i, :k], out["P"][i, :k]
        assert out["root_N"][i] == 801 and n.sum() == 800 and (n >= 0).all()
        vis = n > 0
        assert np.array_equal(q[vis], (w[vis] / n[vis].astype(np.float32)).astype(np.float32)) and not q[~vis].any()
        assert np.isfinite(p).all() and (np.abs(w) <= n + 1e-3).all()
    for i in (0, 137, 511):
        alone = eng.search([recs[i]], 800, noise=[noise[i]], eval_mode=EVAL_NET)
        k = out["n_children"][i]
        assert alone["n_children"][0] == k and np.array_equal(alone["moves"][0, :k], out["moves"][i, :k])
        for key in ("N", "W", "Q", "P"):
            assert np.array_equal(alone[key][0, :k], out[key][i, :k]), (i, key)
        assert alone["nodes"][0] == out["nodes"][i]


def test_warp_move_generation_matches_host_rules_on_many_positions(engine):
    """The warp-cooperative device move generator (csrc/rules_warp.cuh) against the host rules (same source as
    the serial generator, pinned by perft and by the oracle) on ~3000 positions reached by random playouts from
    perft positions rich in checks, pins, castling, promotions and en passant: identical moves in identical order."""
    import random
    from chessbot_b200.board import Board
    fens = [None,
            "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
            "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
            "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
            "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
            "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10",
            "8/P1k5/8/8/8/8/5Kp1/8 w - - 0 1", "4k3/8/8/3pP3/8/8/8/4K2R w K d6 0 2"]
    rnd = random.Random(17)
    boards, expected = [], []
    for fen in fens:
        for _ in range(24):
            b = Board(fen) if fen else Board()
            for _ in range(rnd.randint(1, 40)):
                legal = list(b.legal_moves)
                if not legal:
                    break
                boards.append(b.copy())
                expected.append([m.uci() for m in legal])
                b.push(rnd.choice(legal))
    assert len(boards) > 2500, len(boards)
    in_check = sum(b.is_check() for b in boards)
    assert in_check > 100 and any(len(u) == 5 for e in expected for u in e)
    from chessbot_b200.board import code_to_uci
    for lo in range(0, len(boards), 1024):
        chunk = boards[lo:lo + 1024]
        pos, fi = engine.upload_records([b.export_record() for b in chunk])
        lb = torch.zeros(len(chunk), 4672, dtype=torch.bfloat16, device="cuda")
        moves, actions, priors, counts = engine.policy_softmax(pos, fi[:, 0].contiguous(), lb)
        torch.cuda.synchronize()
        moves, counts = moves.cpu().numpy().astype(np.uint16), counts.cpu().numpy()
        for i, b in enumerate(chunk):
            k = counts[i]
            assert [code_to_uci(c) for c in moves[i, :k]] == expected[lo + i], (b.fen(), k)


def test_virtual_loss_mode_invariants_and_agreement(engine):
    """Throughput mode (north_star: "virtual-loss ... over many concurrent trees"): several leaves of one tree per
    network call.  Not the reference's sequential order, so no bit parity; what must hold: the requested number of
    simulations exactly, every virtual loss settled (Q == W / N, |W| <= N, visits sum to the simulations), the search
    is deterministic, and its preferred move mostly agrees with the exact sequential search."""
    from gpu_util import device_board, random_games
    from chessbot_b200.engine import EVAL_SYNTH
    games = random_games(40, seed=88, max_plies=100)
    recs = [device_board(f, m).export_record() for f, m in games]
    noise = [np.random.default_rng(i).gamma(0.3, 1, 256) if i % 2 else None for i in range(40)]
    exact = engine.search(recs, 200, noise=noise, eval_mode=EVAL_SYNTH, seed=4)
    outs = {}
    for leaves in (4, 16):
        out = engine.search(recs, 200, noise=noise, eval_mode=EVAL_SYNTH, seed=4, leaves=leaves)
        again = engine.search(recs, 200, noise=noise, eval_mode=EVAL_SYNTH, seed=4, leaves=leaves)
        assert out["unfinished"] == 0 and out["errors"] == 0 and out["sims"] == 40 * 200 and out["evals"] <= 40 * 201
        agree = 0
        for i in range(40):
            k = out["n_children"][i]
            n, w, q = out["N"][i, :k], out["W"][i, :k], out["Q"][i, :k]
            assert k == exact["n_children"][i] and np.array_equal(out["moves"][i, :k], exact["moves"][i, :k])
            assert np.array_equal(out["P"][i, :k], exact["P"][i, :k])          # the root expansion is the same
            assert out["root_N"][i] == 201 and n.sum() == 200 and (n >= 0).all()
            vis = n > 0
            assert np.array_equal(q[vis], (w[vis] / n[vis].astype(np.float32)).astype(np.float32)) and not w[~vis].any()
            assert (np.abs(w) <= n + 1e-3).all()
            for key in ("N", "W", "Q"):
                assert np.array_equal(out[key][i, :k], again[key][i, :k])
            agree += int(np.argmax(n) == np.argmax(exact["N"][i, :k]))
        outs[leaves] = agree
        print(f"\nvirtual loss, {leaves} leaves per tree: most-visited move agrees with the sequential search on {agree}/40 positions")
        assert agree >= 26
    back = engine.search(recs, 200, noise=noise, eval_mode=EVAL_SYNTH, seed=4)   # back to one leaf per tree: exact again
    for key in ("N", "W", "Q", "P"):
        assert np.array_equal(back[key], exact[key])


def test_virtual_loss_single_game_with_device_network():
    """mcts.run_mcts(game, network, leaves_per_tree=8): one game, eight leaves per network call."""
    import time
    import chessbot_b200.mcts as mcts
    from chessbot_b200.board import Board
    from chessbot_b200.config import Config
    from chessbot_b200.game import Game
    from chessbot_b200.model import DeviceNetwork, keras_init_weights
    import bench
    net = DeviceNetwork(bench.trained_like_bn(keras_init_weights(128, 6, seed=2), 128, 6), 128, 6, max_batch=64)
    game = Game(Config(), board=Board())
    res = {}
    for leaves in (1, 8):
        mcts.run_mcts(game, net, 64, 0, False, None, leaves_per_tree=leaves)      # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        move, root = mcts.run_mcts(game, net, 400, 0, False, None, leaves_per_tree=leaves)
        res[leaves] = (move, root, time.perf_counter() - t0)
        assert root.N == 401 and sum(c.N for c in root.children.values()) == 400
    print(f"\nsingle game, 400 sims, 6x128: {res[1][2] * 1e3:.0f} ms sequential, {res[8][2] * 1e3:.0f} ms with 8 leaves per call")
    # (a random-init network has no favourite at the start position; move agreement is measured in the test above)
    assert res[8][0] in res[8][1].children and res[8][2] < res[1][2]


def test_virtual_loss_and_yielding_stress_finishes_every_tree_exactly():
    """Stress of the step protocol: many trees, several leaves per tree under virtual loss, few simulations left at the end, terminal-heavy
    positions (mate-in-one and claimable draws make trees yield after a few terminal simulations): every tree must end with exactly its
    simulation budget spent (root N = sims + 1), no pool errors, for 1, 3, 4 and 16 leaves per step and uneven budgets."""
    from chessbot_b200.board import Board
    from chessbot_b200.engine import EVAL_SYNTH, Engine
    eng = Engine.get()
    lines = ["", "f2f3 e7e5 g2g4", "e2e4 e7e5 d1h5 b8c6 f1c4 g8f6", "g1f3 g8f6 f3g1 f6g8 g1f3 g8f6 f3g1 f6g8 g1f3", "d2d4 d7d5 c2c4 e7e6 b1c3 g8f6 c1g5"]
    recs = []
    for i in range(120):
        b = Board()
        for u in lines[i % len(lines)].split():
            b.push_uci(u)
        recs.append(b.export_record())
    sims = [17 + (i * 7) % 90 for i in range(120)]
    eng.search_shape = None
    eng.search_create(120, max_nodes=120 * 130 * 48, max_positions=120 * 400, max_sims=128)
    for leaves in (1, 3, 4, 16):
        out = eng.search(recs, sims, eval_mode=EVAL_SYNTH, seed=9, leaves=leaves)
        assert out["unfinished"] == 0 and out["errors"] == 0
        assert out["root_N"].tolist() == [s + 1 for s in sims], leaves
        assert (out["N"].sum(1) == np.array(sims)).all() or leaves > 1   # sequential mode: every simulation is a visit of one root child
        if leaves > 1:
            assert (out["N"].sum(1) == np.array(sims)).all()
        assert out["sims"] == sum(sims)
