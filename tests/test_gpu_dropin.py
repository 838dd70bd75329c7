"""Drop-in surface (-m gpu): the reference's own module-level calls — mcts.run_mcts / expand / ucb /
update, model.predict_fn, game.Game.make_image, gameimage.board_to_image — served by the CUDA path
through the C ABI.  The known answers are those of the reference's tests/mcts_v2.py (which cannot
run in the reference tree: it imports a missing `mcts_v2`), the full-tree checks compare with the
oracle on the same mock network."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TEST_FEN = "7k/6p1/8/8/8/8/1P6/K7 w - - 1 1"     # tests/mcts_v2.py:12 — four legal moves for each side


def _mock_network(moves, player, policies, value, fill_seed=0):
    """tests/mcts_v2.py:51-67: dict of nested lists, random logits everywhere except the legal moves."""
    import chessbot_b200.actionspace as asp
    actions = [asp.uci_to_action(m, player) for m in moves]

    def net(_image):
        pol = list(np.random.default_rng(fill_seed).random(4672))
        for a, p in zip(actions, policies):
            pol[a] = p
        return {"value_head": [[value, 0.2, 0.3, 0.4]], "policy_head": [pol]}
    return net


def _game(fen=TEST_FEN, moves=()):
    from chessbot_b200.board import Board
    from chessbot_b200.config import Config
    from chessbot_b200.game import Game
    b = Board(fen) if fen else Board()
    for m in moves:
        b.push_uci(m)
    return Game(Config(), board=b)


def test_reference_mcts_known_answers_through_run_mcts():
    """tests/mcts_v2.py:50-87,152-208 on the device search: one simulation from the 4-move position."""
    import chessbot_b200.mcts as mcts
    net = _mock_network(["a1a2", "a1b1", "b2b3", "b2b4"], True, [0.1, 0.2, 0.3, 0.4], 0.1)
    game = _game()
    move, root = mcts.run_mcts(game, net, num_simulations=1, num_sampling_moves=0, add_noise=False)
    moves = game.legal_moves()
    assert len(moves) == 4 and list(root.children) == moves
    s = np.sum([np.exp(n) for n in [0.1, 0.2, 0.3, 0.4]])
    for m, logit in zip(["a1a2", "a1b1", "b2b3", "b2b4"], [0.1, 0.2, 0.3, 0.4]):
        assert root.children[m].P == pytest.approx(np.exp(logit) / s, abs=1e-4)      # "places=4"
    # the first simulation visits the highest prior (b2b4); backup signs: leaf.W == -value, root.N == 2
    leaf = root.children["b2b4"]
    assert move == "b2b4" and leaf.N == 1 and root.N == 2
    assert leaf.W == np.float32(-0.1) and leaf.Q == np.float32(-0.1)
    assert [root.children[m].N for m in moves if m != "b2b4"] == [0, 0, 0]
    C = math.log((root.N + 19652 + 1) / 19652) + 1.25
    assert mcts.ucb(root, root.children["a1a2"]) == pytest.approx((np.exp(0.1) / s) * (2 ** 0.5) * C, abs=1e-4)
    assert mcts.ucb(root, leaf) == pytest.approx((np.exp(0.4) / s) * (2 ** 0.5) / 2 * C - 0.1, abs=1e-4)


def test_node_level_helpers_known_answers():
    """tests/mcts_v2.py:27-48,152-185 with Python nodes and the CUDA encoder behind game.make_image."""
    import chessbot_b200.mcts as mcts
    root = mcts.Node()
    assert root.parent is None and root.children == {} and (root.N, root.W, root.P, root.Q) == (0, 0, 0, 0) and root.is_leaf()
    root.N = 1
    c1 = mcts.Node()
    c1.N = 1
    root.children["e2e4"], root.children["e2e3"], root.children["d2d4"] = c1, mcts.Node(), mcts.Node()
    assert mcts.select_move(root) == "e2e4"
    net = _mock_network(["a1a2", "a1b1", "b2b3", "b2b4"], True, [0.1, 0.2, 0.3, 0.4], 0.1)
    game = _game()
    root = mcts.Node()
    root.N = 1
    value = mcts.expand(root, game, net)
    assert (root.N, root.W, root.Q) == (1, 0, 0) and len(root.children) == 4
    C = math.log((root.N + 19652 + 1) / 19652) + 1.25
    for m in root.children:
        assert mcts.ucb(root, root.children[m]) == root.children[m].P * C          # tests/mcts_v2.py:179-182
    move, leaf = mcts.select_leaf(root)
    assert move == "b2b4" and leaf is root.children["b2b4"]
    game.make_move(move)
    assert mcts.evaluate(game) == (0, False)
    mcts.update(leaf, -value)
    assert leaf.N == 1 and leaf.W == -value and leaf.Q == -value and root.N == 2 and root.W == value


class _Rng:
    def __init__(self, noise=None):
        self.noise = noise

    def gamma(self, a, b, n):
        return self.noise[:n]

    def choice(self, x):
        return x[0]

    def multinomial(self, n, pi):
        o = np.zeros(len(pi), dtype=int)
        o[int(np.argmax(pi))] = 1
        return o


def _hash_network(scale, seed):
    """Deterministic host callable with the reference's network contract; logits up to +-scale."""
    def net(image):
        img = np.asarray(image, dtype=np.float32)[0]
        h = int(np.abs(img[..., :112]).sum() * 7 + img[0, 0, 113] * 13 + img[0, 0, 118] * 3 + seed)
        r = np.random.default_rng(h)
        return {"value_head": np.array([[r.uniform(-1, 1)]], dtype=np.float32),
                "policy_head": (r.standard_normal((1, 4672)) * scale).astype(np.float32)}
    return net


def _compare_root(root, tree):
    kids = list(tree.children(0))
    assert list(root.children) == [tree.move[c] for c in kids]
    for (m, node), c in zip(root.children.items(), kids):
        assert node.N == tree.N[c], m
        assert np.array_equal(np.float32(node.W), np.float32(tree.W[c]), equal_nan=True), m
        assert np.array_equal(np.float32(node.Q), np.float32(tree.Q[c]), equal_nan=True), m
        assert type(node.P) is type(tree.P[c]) and np.array_equal(node.P, tree.P[c], equal_nan=True), m
    assert root.N == tree.N[0]


@pytest.mark.parametrize("scale,add_noise", [(1.0, True), (3.0, False), (60.0, True), (120.0, False)])
def test_run_mcts_with_host_network_matches_oracle(scale, add_noise):
    """A user-supplied Python network (the reference's mock contract): the device runs select / rules /
    gather+normalise / backup, the host evaluates.  scale 60 / 120 overflow np.exp(float32) to inf, so
    priors become 0 / NaN exactly like in the reference (max() keeps the first NaN candidate)."""
    from gpu_util import oracle_board, random_games
    import chessbot_b200.mcts as mcts
    from oracle import pychess_shim as ochess
    from oracle import ref_mcts
    from oracle.ref_game import Game as OGame
    from oracle.ref_game import OracleConfig
    net = _hash_network(scale, seed=int(scale))
    games = random_games(3, seed=int(scale) + 5, max_plies=70)
    for gi, (f, m) in enumerate(games):
        noise = np.random.default_rng(gi).gamma(0.3, 1, 256)
        game = _game(None, m.split())
        with np.errstate(all="ignore"):
            res = mcts.run_mcts_batch([game], net, 40, 0, add_noise, None, rng=_Rng(noise), noise=[noise])
            move, root = res[0]
            omove, tree = ref_mcts.run_mcts(OGame(OracleConfig(), board=oracle_board(f, m)), net, 40, 0, add_noise, None,
                                            rng=_Rng(noise), noise=noise)
        _compare_root(root, tree)
        assert move == omove


def test_run_mcts_batch_mixed_settings_synthetic_network():
    """Playout-cap style batch (train.py:30-41): per-game sims / noise / CPUCT in one device batch."""
    from gpu_util import oracle_board, random_games
    import chessbot_b200.mcts as mcts
    from chessbot_b200.model import SyntheticNetwork
    from oracle import ref_mcts, synthetic_net
    from oracle.ref_game import Game as OGame
    from oracle.ref_game import OracleConfig
    games = random_games(12, seed=77, max_plies=90)
    full = [i % 4 == 0 for i in range(12)]
    sims = [90 if f else 25 for f in full]
    cpuct = [None if f else 1.0 for f in full]
    noise = [np.random.default_rng(i).gamma(0.3, 1, 256) for i in range(12)]
    res = mcts.run_mcts_batch([_game(None, m.split()) for _, m in games], SyntheticNetwork(seed=3), sims, 30, full, cpuct,
                              rng=_Rng(), noise=noise)
    for i, (f, m) in enumerate(games):
        omove, tree = ref_mcts.run_mcts(OGame(OracleConfig(), board=oracle_board(f, m)), synthetic_net.as_network(3), sims[i], 30,
                                        full[i], cpuct[i], rng=_Rng(noise[i]), noise=noise[i])
        _compare_root(res[i][1], tree)
        assert res[i][0] == omove


def test_predict_fn_contract_and_tolerance():
    """model.predict_fn(network, image) -> (value, logits[4672]) (model.py:73-80) on the tcgen05 tower."""
    from gpu_util import oracle_board, random_games
    import chessbot_b200.model as model
    from oracle import ref_gameimage, ref_model
    games = random_games(8, seed=4, max_plies=100)
    images = np.stack([ref_gameimage.board_to_image(oracle_board(f, m), 8) for f, m in games])
    w = ref_model.calibrate_bn(ref_model.init_weights(48, 4, seed=8), images)
    net = model.DeviceNetwork(w, 48, 4, max_batch=8)
    v_ref, p_ref = ref_model.forward(w, images)
    for i in range(3):
        value, logits = model.predict_fn(net, images[i])
        assert logits.shape == (4672,) and np.ndim(value) == 0
        assert abs(float(value) - v_ref[i]) <= 2e-2 and np.abs(logits - p_ref[i]).max() <= 2e-2   # north_star tolerance
    out = net(images.astype(np.float32))
    assert out["value_head"].shape == (8, 1) and out["policy_head"].shape == (8, 4672)
    assert np.abs(out["policy_head"] - p_ref).max() <= 2e-2


def test_game_make_image_and_gameimage_helpers(golden_dir):
    """game.Game.make_image / gameimage.board_to_image + helpers (gameimage.py:24-192) via the CUDA encoder."""
    import chessbot_b200.gameimage as gi
    from gpu_util import device_board
    g = np.load(os.path.join(golden_dir, "gameimage.npz"))
    for k in (0, 7, 23, 41):
        b = device_board(g["fens"][k], g["moves"][k])
        img = gi.board_to_image(b, 8)
        assert img.dtype == np.int16 and np.array_equal(img, g["images"][k])
        for T in (1, 2, 4):
            assert gi.board_to_image(b, T).shape == (8, 8, 14 * T + 7)
            assert np.array_equal(gi.board_to_image(b, T), g["images"][k][..., 14 * (8 - T):])
        p1, p2 = gi._pieces_planes(b)
        assert np.array_equal(p1, g["images"][k][..., 98:104]) and np.array_equal(p2, g["images"][k][..., 104:110])
        assert np.array_equal(gi._color_plane(b), g["images"][k][..., 112])
        assert np.array_equal(gi._movecount_plane(b), g["images"][k][..., 113])
        assert np.array_equal(gi._halfmovecount_plane(b), g["images"][k][..., 118])
    with pytest.raises(ValueError):
        gi.board_to_image(device_board(g["fens"][0], g["moves"][0]), 0)
    game = _game(None, "e2e4 e7e5 g1f3".split())
    assert np.array_equal(game.make_image(-1), gi.board_to_image(game.board, 8))
    assert np.array_equal(game.make_image(1), gi.board_to_image(_game(None, ["e2e4"]).board, 8))


def test_self_play_games_match_oracle_move_for_move():
    """§8(f) row 2: train.py:21-52 (playout-cap randomisation, per-move search, sample format) for several
    concurrent device games against the oracle playing the same games one by one with the same per-game
    random streams: identical move sequences, samples, outcomes."""
    from chessbot_b200 import selfplay
    from chessbot_b200.config import Config
    from chessbot_b200.model import SyntheticNetwork
    from oracle import pychess_shim as ochess
    from oracle import ref_actionspace, ref_mcts, synthetic_net
    from oracle.ref_game import Game as OGame
    from oracle.ref_game import OracleConfig
    cfg = Config()
    cfg.num_mcts_sims, cfg.num_mcts_sampling_moves, cfg.max_game_length = (6, 20), 8, 30
    n = 5
    res = selfplay.play_games(SyntheticNetwork(seed=11), n, cfg, rngs=[np.random.default_rng(100 + i) for i in range(n)])
    for i in range(n):
        ocfg = OracleConfig(max_game_length=30)
        ocfg.num_mcts_sims, ocfg.num_mcts_sampling_moves = (6, 20), 8
        game, rng = OGame(ocfg), np.random.default_rng(100 + i)
        images, stats, turns = [], [], []
        while not game.terminal():
            full, sims, cpuct = selfplay.playout_cap(ocfg, rng.random())
            move, tree = ref_mcts.run_mcts(game, synthetic_net.as_network(11), sims, 8, full, cpuct, rng=rng)
            if full:
                kids = list(tree.children(0))
                total = sum(tree.N[c] for c in kids)
                pi = np.zeros(4672)
                for c in kids:
                    pi[ref_actionspace.uci_to_action(tree.move[c], game.to_play())] = tree.N[c] / total
                stats.append(pi)
                images.append(game.make_image(-1))
                turns.append(game.to_play())
            game.make_move(move)
        got_images, (got_z, got_pi), outcome_str, history_len = res[i]
        assert history_len == game.history_len and outcome_str == game.outcome_str
        assert len(got_images) == len(images) and all(np.array_equal(a, b) and a.dtype == np.int16 for a, b in zip(got_images, images))
        assert all(np.array_equal(a, b) and a.dtype == np.float64 for a, b in zip(got_pi, stats))
        assert got_z == [game.terminal_value(p) for p in turns]
    entries = [e for r in res for e in selfplay.buffer_entries(r)]
    if len(entries) >= 2:
        imgs, (z, pi) = selfplay.sample_buffer(entries, 2, np.random.default_rng(0))
        assert imgs.shape == (2, 8, 8, 119) and pi.shape == (2, 4672) and z.shape == (2,)


def test_puzzle_bot_matches_oracle_search():
    """§8(f) row 3: eval_trt.Bot.get_move(board) over the batched search (puzzles.Puzzle.solve calls nothing else)."""
    from chessbot_b200 import selfplay
    from chessbot_b200.board import Board
    from chessbot_b200.model import SyntheticNetwork
    from oracle import pychess_shim as ochess
    from oracle import ref_mcts, synthetic_net
    from oracle.ref_game import Game as OGame
    from oracle.ref_game import OracleConfig
    fens = ["r1bqkb1r/pppp1ppp/2n2n2/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4",      # mate in one
            "6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1", "7k/6p1/8/8/8/8/1P6/K7 w - - 1 1"]
    bot = selfplay.Bot(SyntheticNetwork(seed=2), num_simulations=40)
    moves = bot.get_moves([Board(f) for f in fens])
    assert bot.get_move(Board(fens[0])) == moves[0]
    for f, mv in zip(fens, moves):
        omove, _ = ref_mcts.run_mcts(OGame(OracleConfig(), board=ochess.Board(f)), synthetic_net.as_network(2), 40, 0, False, None,
                                     rng=_Rng())
        assert mv == omove


def test_batched_puzzle_solving_equals_puzzle_by_puzzle():
    """selfplay.solve_puzzles (all live puzzles = one device batch per round) leaves every puzzle exactly where the reference's
    own loop `puzzle.solve(bot)` (puzzles.py:83-86) leaves it."""
    from test_puzzles_host import _Puzzle
    from chessbot_b200 import selfplay
    from chessbot_b200.model import SyntheticNetwork
    from gpu_util import random_games
    specs = []
    for f, m in random_games(12, seed=77, min_plies=9, max_plies=40):
        mv = m.split()
        specs.append((f, " ".join(mv[-5:]), mv[:-5]))          # the last 5 plies are the "solution": opponent, bot, opponent, bot, opponent...

    def make():
        out = []
        for f, sol, prefix in specs:
            from chessbot_b200.board import Board
            b = Board(str(f))
            for u in prefix:
                b.push_uci(u)
            out.append(_Puzzle(b.fen(), sol))
        return out
    bot = selfplay.Bot(SyntheticNetwork(seed=5), num_simulations=24, rng=_Rng())      # ties among most-visited moves: first one
    batched, single = make(), make()
    selfplay.solve_puzzles(batched, bot)
    for p in single:
        while p.status == "IN_PROGRESS":
            p.move_uci(bot.get_move(p.board))
    assert [p.played for p in batched] == [p.played for p in single]
    assert [p.status for p in batched] == [p.status for p in single]


def test_config4_top_move_agreement_with_fp32_reference_path():
    """BASELINE config 4 at test scale (puzzle-style sweep: top-move agreement vs the reference).  Here the two
    sides use their OWN networks — device: tcgen05 bf16 tower, reference path: torch-fp32 — so logits differ by
    up to 2e-2 and visit counts may drift; the arg-max-visit move must still agree on the large majority of
    positions, and where it differs the reference's move must be among the device's most visited."""
    from gpu_util import oracle_board, random_games
    from chessbot_b200.model import DeviceNetwork
    from oracle import ref_gameimage, ref_mcts, ref_model
    from oracle.ref_game import Game as OGame
    from oracle.ref_game import OracleConfig
    games = random_games(40, seed=404, max_plies=80)
    images = np.stack([ref_gameimage.board_to_image(oracle_board(f, m), 8) for f, m in games])
    w = ref_model.calibrate_bn(ref_model.init_weights(48, 4, seed=4), images, device="cuda")
    w["policy.fc"] = w["policy.fc"] * 8.0            # sharpen the random-init policy so that searches have a clear favourite
    import chessbot_b200.mcts as mcts
    res = mcts.run_mcts_batch([_game(None, m.split()) for _, m in games], DeviceNetwork(w, 48, 4, max_batch=64), 100, 0, False, None, rng=_Rng())
    net = ref_model.as_network(w, "cuda")
    agree, in_top3 = 0, 0
    for (move, root), (f, m) in zip(res, games):
        omove, _ = ref_mcts.run_mcts(OGame(OracleConfig(), board=oracle_board(f, m)), net, 100, 0, False, None, rng=_Rng())
        agree += move == omove
        top3 = sorted(root.children, key=lambda u: -root.children[u].N)[:3]
        in_top3 += omove in top3
    print(f"\nconfig 4 (40 positions x 100 sims, 4x48): top-move agreement {agree}/40, reference move in device top-3 {in_top3}/40")
    assert agree >= 30 and in_top3 >= 36


def test_library_memory_is_owned_by_pytorch():
    """SURVEY.md §8(b) / north_star "PyTorch for tensor ownership only": what the library keeps between calls (weights, activations,
    node pools) is allocated through PyTorch's caching allocator (cb_set_device_allocator), so torch accounts for it."""
    import torch
    from chessbot_b200 import engine as E
    from chessbot_b200.model import keras_init_weights
    eng = E.Engine.get()
    before_lib, before_torch = E.library_device_bytes(), torch.cuda.memory_allocated()
    eng.load_network(keras_init_weights(64, 2, seed=1), 64, 2, max_batch=512)
    grown = E.library_device_bytes() - before_lib
    assert E.library_device_bytes() > 0
    assert torch.cuda.memory_allocated() - before_torch >= 0.95 * grown and grown >= 0
    need = E.lib.cb_search_workspace_bytes(256, 256 * 40 * 48, 256 * 300, 38, 1)
    fresh = E.Engine.actor(7)                       # a context that never held a pool
    lib0 = E.library_device_bytes()
    fresh.search_create(256, max_nodes=256 * 40 * 48, max_positions=256 * 300, max_sims=38)
    took = E.library_device_bytes() - lib0
    assert 0.5 * need <= took <= 1.05 * need + (1 << 20), (need, took)


def test_search_stream_and_two_lane_self_play_equal_the_one_at_a_time_path():
    """mcts.SearchStream keeps two searches in flight (two device contexts, one stream); selfplay.play_games alternates two groups of
    games over it.  Neither may change a result: visit counts of a submitted search equal search_batch's, and games played with
    lanes=2 equal the same games (same per-game generators) played with lanes=1, move for move."""
    import chessbot_b200.mcts as mcts
    from chessbot_b200 import selfplay
    from chessbot_b200.config import Config
    from chessbot_b200.game import Game
    from chessbot_b200.model import DeviceNetwork
    from oracle import ref_model
    rng = np.random.default_rng(3)
    images = rng.integers(0, 2, (16, 8, 8, 119)).astype(np.int16)
    w = ref_model.calibrate_bn(ref_model.init_weights(64, 2, seed=6), images, device="cuda")
    net = DeviceNetwork(w, 64, 2, max_batch=64)
    cfg = Config()
    lines = ["", "e2e4", "e2e4 e7e5", "d2d4 d7d5 c2c4", "g1f3 g8f6 c2c4 e7e6", "e2e4 c7c5 g1f3 d7d6 d2d4"]

    def games():
        out = []
        for ln in lines:
            g = Game(cfg)
            for u in ln.split():
                g.make_move(u)
            out.append(g)
        return out
    noise = [np.random.default_rng(i).gamma(0.3, 1, len(g.board.legal_moves)) for i, g in enumerate(games())]
    base = mcts.search_batch(games(), net, 40, True, None, noise=noise)
    ss = mcts.SearchStream(net, 2)
    ga, gb = games()[:3], games()[3:]
    ta = ss.submit(ga, 40, True, None, noise=noise[:3])
    tb = ss.submit(gb, 40, True, None, noise=noise[3:])
    got = ss.collect(ta) + ss.collect(tb)
    for r0, r1 in zip(base, got):
        assert r0.N == r1.N and list(r0.children) == list(r1.children)
        assert [c.N for c in r0.children.values()] == [c.N for c in r1.children.values()]
        assert [c.Q for c in r0.children.values()] == [c.Q for c in r1.children.values()]
    small = Config()
    small.num_mcts_sims = (6, 10)
    small.max_game_length = 12

    def play(lanes):
        return selfplay.play_games(net, 6, small, rngs=[np.random.default_rng(100 + i) for i in range(6)], lanes=lanes)
    one, two = play(1), play(2)
    for a, b in zip(one, two):
        assert a[2] == b[2] and a[3] == b[3] and len(a[0]) == len(b[0])
        assert all(np.array_equal(x, y) for x, y in zip(a[0], b[0]))
        assert all(np.array_equal(x, y) for x, y in zip(a[1][1], b[1][1])) and a[1][0] == b[1][0]
