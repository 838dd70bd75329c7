"""ORACLE (test infrastructure). Closed-form restatement of actionspace.py.

Reference: `uci_to_action(uci, player)` actionspace.py:116-180; vocabulary
actionspace.py:4-42.  index = (file'*8 + rank')*73 + plane with (file', rank')
seen from the mover (Black: both reversed, actionspace.py:126-127).
Pinned against the reference file imported verbatim for all 2*64*63*5 inputs
(scripts/make_golden.py) and against tests/actionspace.py's 200 known answers.
"""
import numpy as np

NUM_PLANES = 73
NUM_ACTIONS = 8 * 8 * NUM_PLANES
_DIRS = ["N", "NE", "E", "SE", "S", "SW", "W", "NW"]
_DIR_VEC = {"N": (0, 1), "NE": (1, 1), "E": (1, 0), "SE": (1, -1),
            "S": (0, -1), "SW": (-1, -1), "W": (-1, 0), "NW": (-1, 1)}
# knight plane order 56..63 as (dx, dy): actionspace.py:14-19
_KNIGHT = [(1, 2), (-1, 2), (2, 1), (2, -1), (1, -2), (-1, -2), (-2, 1), (-2, -1)]
_KNIGHT_NAMES = ["K,N,E", "K,N,W", "K,E,N", "K,E,S", "K,S,E", "K,S,W", "K,W,N", "K,W,S"]
_UNDER = {"r": 0, "b": 1, "n": 2}

moves = np.array([f"Q,{d},{n}" for n in range(1, 8) for d in _DIRS] + _KNIGHT_NAMES
                 + [f"P,{d},{p}" for d in ("N", "NE", "NW") for p in ("R", "B", "K")], dtype=object)
moves_dict = {m: i for i, m in enumerate(moves)}
action_space = np.tile(moves, 64)
action_space_indices = np.arange(NUM_ACTIONS).reshape(8, 8, NUM_PLANES)


def move_plane(dx: int, dy: int, promotion: str | None):
    """Plane 0..72 for a displacement seen from the mover, or None."""
    if (dx, dy) in _KNIGHT:
        return 56 + _KNIGHT.index((dx, dy))
    if promotion and dy == 1 and dx in (0, 1, -1):
        if promotion == "q":
            return {0: 0, 1: 1, -1: 7}[dx]
        return 64 + {0: 0, 1: 1, -1: 2}[dx] * 3 + _UNDER[promotion]
    if (dx == 0 or dy == 0 or abs(dx) == abs(dy)) and (dx or dy):
        dist = max(abs(dx), abs(dy))
        sx, sy = (dx > 0) - (dx < 0), (dy > 0) - (dy < 0)
        d = [_DIR_VEC[k] for k in _DIRS].index((sx, sy))
        return 8 * (dist - 1) + d
    return None


def uci_to_action(uci: str, player: bool):
    fc, fr = "abcdefgh".index(uci[0]), "12345678".index(uci[1])
    tc, tr = "abcdefgh".index(uci[2]), "12345678".index(uci[3])
    if not player:
        fc, fr, tc, tr = 7 - fc, 7 - fr, 7 - tc, 7 - tr
    promo = uci[4] if len(uci) == 5 else None
    plane = move_plane(tc - fc, tr - fr, promo)
    if plane is None:
        return None
    return (fc * 8 + fr) * NUM_PLANES + plane


def move_to_action(from_sq: int, to_sq: int, promotion_piece: int, white_to_move: bool):
    """Same index from square numbers (a1=0) and python-chess piece type (0/None, 2..5)."""
    promo = {0: None, None: None, 2: "n", 3: "b", 4: "r", 5: "q"}[promotion_piece]
    uci = "abcdefgh"[from_sq & 7] + "12345678"[from_sq >> 3] + "abcdefgh"[to_sq & 7] + "12345678"[to_sq >> 3]
    return uci_to_action(uci + (promo or ""), white_to_move)


def uci_to_actionstr(uci: str, player: bool):
    return action_space[uci_to_action(uci, player)]


def action_idx_to_str(idx: int) -> str:
    return action_space[idx]
