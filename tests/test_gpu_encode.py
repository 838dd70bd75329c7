"""K1 parity (-m gpu): CUDA board->planes through the C ABI against the reference's own images
(tests/golden/gameimage.npz, produced by gameimage.py verbatim) and against the oracle on seeded
random playouts.  Bar: bit-exact integers."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from chessbot_b200.engine import Engine
    return Engine.get()


def _encode(engine, boards):
    from chessbot_b200.engine import act_to_nhwc
    pos, fi = engine.upload_records([b.export_record() for b in boards])
    act, i16 = engine.encode(pos, fi, want_act=True, want_i16=True)
    return i16.cpu().numpy(), act_to_nhwc(act, len(boards), 128).float().cpu().numpy()


def _check_bf16_planes(planes, images):
    # channels 0..118 hold the image; 119/120 the bf16 remainders of move count / halfmove clock
    ref = images.astype(np.float32)
    got = planes[..., :119].copy()
    got[..., 113] += planes[..., 119]
    got[..., 118] += planes[..., 120]
    assert np.array_equal(got, ref)
    assert not planes[..., 121:].any()
    # what stays in channel 113 / 118 must be exactly representable in bf16
    from gpu_util import bf16_round
    assert np.array_equal(bf16_round(planes[..., 113]), planes[..., 113])


def test_golden_images_bit_exact(engine, golden_dir):
    from gpu_util import device_board
    g = np.load(os.path.join(golden_dir, "gameimage.npz"))
    boards = [device_board(f, m) for f, m in zip(g["fens"], g["moves"])]
    i16, planes = _encode(engine, boards)
    assert i16.dtype == np.int16 and i16.shape == (50, 8, 8, 119)
    assert np.array_equal(i16, g["images"])
    _check_bf16_planes(planes, g["images"])


def test_random_playouts_against_oracle(engine):
    from gpu_util import device_board, oracle_board, random_games
    from oracle import ref_gameimage
    games = random_games(257, seed=11, max_plies=200, back_bias=0.3)  # odd count: half-empty last board pair
    boards = [device_board(f, m) for f, m in games]
    i16, planes = _encode(engine, boards)
    ref = np.stack([ref_gameimage.board_to_image(oracle_board(f, m), 8) for f, m in games])
    assert np.array_equal(i16, ref)
    _check_bf16_planes(planes, ref)
    assert (ref[:, 0, 0, 12:112:14] > 0).sum() > 10  # repetition planes were exercised


def test_long_game_move_count_beyond_bf16(engine):
    # move count > 256 is not representable in bf16: the remainder channel must make it exact
    from gpu_util import device_board, oracle_board
    from oracle import ref_gameimage
    shuffle = "g1f3 g8f6 f3g1 f6g8 b1c3 b8c6 c3b1 c6b8".split()
    adv = ["a2a3", "a7a6", "a3a4", "a6a5", "b2b3", "b7b6", "b3b4", "b6b5", "c2c3", "c7c6", "c3c4", "c6c5", "d2d3", "d7d6",
           "d3d4", "d6d5", "e2e3", "e7e6", "e3e4", "e6e5", "f2f3", "f7f6", "f3f4", "f6f5", "g2g3", "g7g6", "g3g4", "g6g5",
           "h2h3", "h7h6", "h3h4", "h6h5"]
    moves = []
    k = 0
    while len(moves) < 330 and k + 1 < len(adv):
        moves += shuffle[:4] + shuffle[4:] + adv[k:k + 2]   # 8 reversible plies then two pawn moves
        k += 2
    ob = oracle_board("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", "")
    ok_moves = []
    for m in moves:
        if ob.is_game_over(claim_draw=True):
            break
        try:
            ob.push_uci(m)
        except ValueError:
            break
        ok_moves.append(m)
    if ob.is_game_over(claim_draw=True):
        ob.pop()
        ok_moves.pop()
    assert len(ok_moves) > 256
    b = device_board("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", " ".join(ok_moves))
    i16, planes = _encode(engine, [b])
    ref = ref_gameimage.board_to_image(ob, 8)[None]
    assert ref[0, 0, 0, 113] == len(ok_moves)
    assert np.array_equal(i16, ref)
    _check_bf16_planes(planes, ref)


def test_fen_root_has_single_frame(engine):
    from gpu_util import device_board
    b = device_board("r2qk2r/pppb1ppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPPB1PPP/R2QK2R w KQkq - 4 7", "")
    i16, _ = _encode(engine, [b])
    assert not i16[0, :, :, :98].any() and i16[0, :, :, 98:110].any()
    assert (i16[0, :, :, 113] == 0).all() and (i16[0, :, :, 118] == 4).all() and (i16[0, :, :, 114:118] == 1).all()
