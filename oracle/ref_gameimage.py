"""ORACLE (test infrastructure). Closed-form restatement of gameimage.py.

Reference: `board_to_image(board, T)` gameimage.py:149-192 and helpers :24-146.
Square -> (row, col): White to move row=7-rank, col=file; Black to move the
board is mirrored and colour-swapped (gameimage.py:37-39) => row=rank,
col=7-file, "player 1" = the side to move OF THAT FRAME (gameimage.py:170,178).
Works on any python-chess-compatible board (oracle.pychess_shim.Board).
Pinned against the reference file imported verbatim (scripts/make_golden.py)
and against the reference's tests/gameimage.py.
"""
import numpy as np

M = 14
L = 7
# plane offset within a colour block: P,N,B,R,K,Q (gameimage.py:10-17); python-chess types 1..6
_PLANE_OF_TYPE = {1: 0, 2: 1, 3: 2, 4: 3, 6: 4, 5: 5}


def _scatter(image, base, mask, white_to_move):
    while mask:
        low = mask & -mask
        sq = low.bit_length() - 1
        mask ^= low
        rank, file = sq >> 3, sq & 7
        if white_to_move:
            image[7 - rank, file, base] = 1
        else:
            image[rank, 7 - file, base] = 1


def frame_planes(image, board, slot):
    """Write the 14 planes of one history frame at channel 14*slot."""
    w = bool(board.turn)
    base = slot * M
    for pt, off in _PLANE_OF_TYPE.items():
        _scatter(image, base + off, int(board.pieces_mask(pt, w)), w)
        _scatter(image, base + 6 + off, int(board.pieces_mask(pt, not w)), w)
    if board.is_repetition(2):
        image[:, :, base + 12] = 1
    if board.is_repetition(3):
        image[:, :, base + 13] = 1


def board_to_image(board, T: int = 8) -> np.ndarray:
    if T < 1:
        raise ValueError("T must be greater than 0")
    image = np.zeros((8, 8, M * T + L), dtype=np.int16)
    tmp = board.copy()
    for t in range(T):
        frame_planes(image, tmp, T - 1 - t)
        if len(tmp.move_stack) > 0:
            tmp.pop()
        else:
            break
    w = bool(board.turn)
    c = M * T
    image[:, :, c] = 1 if w else 0
    image[:, :, c + 1] = len(board.move_stack)
    image[:, :, c + 2] = int(board.has_kingside_castling_rights(w))
    image[:, :, c + 3] = int(board.has_queenside_castling_rights(w))
    image[:, :, c + 4] = int(board.has_kingside_castling_rights(not w))
    image[:, :, c + 5] = int(board.has_queenside_castling_rights(not w))
    image[:, :, c + 6] = board.halfmove_clock
    return image
