"""Drop-in for the reference's actionspace.py: the 73-plane x 64-square move vocabulary
(actionspace.py:4-42) and uci <-> action index (actionspace.py:45-194).  The index itself is computed
by the C ABI (cb_uci_to_action; the same closed form the device kernels use, csrc/rules.cuh)."""
import numpy as np

from ._lib import lib

moves = []
for num_squares in range(1, 8):
    for direction in ["N", "NE", "E", "SE", "S", "SW", "W", "NW"]:
        moves.append(f"Q,{direction},{num_squares}")
for l, direction_long in enumerate(["N", "E", "S", "W"]):
    for s, direction_short in enumerate(["N", "E", "S", "W"]):
        if direction_long != direction_short and abs(l - s) % 2 == 1:
            moves.append(f"K,{direction_long},{direction_short}")
for direction in ["N", "NE", "NW"]:
    for piece in ["R", "B", "K"]:
        moves.append(f"P,{direction},{piece}")
moves = np.array(moves, dtype=object)
moves_dict = {move: i for i, move in enumerate(moves)}
action_space = np.empty((8, 8, len(moves)), dtype=moves.dtype)
action_space[:, :, :] = moves
action_space = action_space.flatten()
action_space_indices = np.arange(len(action_space)).reshape((8, 8, len(moves)))

_columns_str = ["a", "b", "c", "d", "e", "f", "g", "h"]
_rows_str = ["1", "2", "3", "4", "5", "6", "7", "8"]
promotions_dict = {"r": "R", "b": "B", "n": "K", "q": "Q"}


def get_board_indices() -> np.ndarray:
    return action_space_indices


def get_action_idx(square: tuple, move: str) -> int:
    # same checks, in the same order, as actionspace.py:62-72 (the membership test is against the
    # dictionary's integer values, so a move *string* is rejected exactly like in the reference)
    if len(square) != 2:
        raise ValueError("Square must be a tuple of length 2")
    if square[0] not in range(8) or square[1] not in range(8):
        raise ValueError("Square must be a tuple of length 2 with values in range(8)")
    if move not in moves_dict.values():
        raise ValueError("Move must be a valid move")
    col, row = square
    return action_space_indices[col, row, moves_dict[move]]


def action_idx_to_str(idx: int) -> str:
    return action_space[idx]


def uci_to_action(uci: str, player: bool):
    """Index 0..4671 of a UCI move seen from `player` (True = White), None if the displacement is not in
    the vocabulary (actionspace.py:116-180).  Malformed squares raise ValueError like the reference's
    list.index()."""
    if len(uci) < 4 or uci[0] not in _columns_str or uci[1] not in _rows_str or uci[2] not in _columns_str \
            or uci[3] not in _rows_str:
        raise ValueError(f"{uci!r} is not in list")
    text = uci[:4] + (uci[4] if len(uci) == 5 and uci[4] in "qrbn" else "")
    a = lib.cb_uci_to_action(text.encode(), int(bool(player)))
    if len(uci) == 5 and uci[4] not in "qrbn" and a >= 0:
        # the reference looks the letter up in promotions_dict only on the promotion path
        dx = _columns_str.index(uci[2]) - _columns_str.index(uci[0])
        dy = (_rows_str.index(uci[3]) - _rows_str.index(uci[1])) * (1 if player else -1)
        if dy == 1 and abs(dx) <= 1 and abs(dx * dy) != 2:
            raise KeyError(uci[4])
    return None if a < 0 else a


def uci_to_actionstr(uci: str, player: bool):
    return action_space[uci_to_action(uci, player)]
