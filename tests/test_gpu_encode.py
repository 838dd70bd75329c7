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
    import random
    from gpu_util import device_board
    from oracle import pychess_shim as ochess
    from oracle import ref_gameimage
    rnd = random.Random(5)
    ob, moves = None, []
    while len(moves) < 301:
        ob, moves = ochess.Board(), []
        while len(moves) < 301 and not ob.is_game_over(claim_draw=False):
            legal = list(ob.legal_moves)
            quiet = [m for m in legal if not ob.is_zeroing(m)]
            m = rnd.choice(quiet if quiet and len(moves) % 40 != 39 else legal)  # a pawn move / capture now and then
            ob.push(m)
            moves.append(m.uci())
    b = device_board(ochess.STARTING_FEN, " ".join(moves))
    i16, planes = _encode(engine, [b])
    ref = ref_gameimage.board_to_image(ob, 8)[None]
    assert ref[0, 0, 0, 113] == 301
    assert np.array_equal(i16, ref)
    _check_bf16_planes(planes, ref)
    assert planes[0, 0, 0, 113] == 300 and planes[0, 0, 0, 119] == 1


def test_fen_root_has_single_frame(engine):
    from gpu_util import device_board
    b = device_board("r2qk2r/pppb1ppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPPB1PPP/R2QK2R w KQkq - 4 7", "")
    i16, _ = _encode(engine, [b])
    assert not i16[0, :, :, :98].any() and i16[0, :, :, 98:110].any()
    assert (i16[0, :, :, 113] == 0).all() and (i16[0, :, :, 118] == 4).all() and (i16[0, :, :, 114:118] == 1).all()
