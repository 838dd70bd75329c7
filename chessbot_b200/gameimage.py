"""Drop-in for the reference's gameimage.py: board -> int16 [8, 8, 14*T + 7] planes
(gameimage.py:149-192).  The planes are produced by the CUDA encoder (csrc/encode.cu) through
cb_encode; this module only adapts the board and slices frames for T < 8."""
import numpy as np

N = 8
M = 6 + 6 + 2


def as_device_board(board):
    """chessbot_b200.board.Board for any python-chess-compatible board (replays its move stack)."""
    from .board import Board
    if isinstance(board, Board):
        return board
    root = board.root() if hasattr(board, "root") else None
    if root is None:
        tmp = board.copy()
        moves = []
        while len(tmp.move_stack) > 0:
            moves.append(tmp.pop().uci())
        root, moves = tmp, moves[::-1]
    else:
        moves = [m.uci() for m in board.move_stack]
    b = Board(root.fen())
    for u in moves:
        b.push_uci(u)
    return b


def _encode(boards):
    from .engine import Engine
    eng = Engine.get()
    pos, fi = eng.upload_records([as_device_board(b).export_record() for b in boards])
    _, i16 = eng.encode(pos, fi, want_act=False, want_i16=True)
    return i16.cpu().numpy()


def boards_to_images(boards, T: int = 8) -> np.ndarray:
    """Batched board_to_image: int16 [n, 8, 8, 14*T + 7]."""
    if T < 1:
        raise ValueError("T must be greater than 0")
    if T > 8:
        raise NotImplementedError("the device encoder holds 8 history frames (config.T = 8)")
    img = _encode(boards)
    return np.ascontiguousarray(img[..., 14 * (8 - T):])


def board_to_image(board, T: int = 8) -> np.ndarray:
    return boards_to_images([board], T)[0]


def _pieces_planes(board):
    img = board_to_image(board, 1)
    return img[:, :, 0:6].copy(), img[:, :, 6:12].copy()


def _repetitions_planes(board):
    return board_to_image(board, 1)[:, :, 12:14].copy()


def _color_plane(board):
    return board_to_image(board, 1)[:, :, 14].copy()


def _movecount_plane(board):
    return board_to_image(board, 1)[:, :, 15].copy()


def _castlingrights_planes(board):
    img = board_to_image(board, 1)
    p1, p2 = img[:, :, 16:18].copy(), img[:, :, 18:20].copy()
    return p1, p2


def _halfmovecount_plane(board):
    return board_to_image(board, 1)[:, :, 20].copy()
