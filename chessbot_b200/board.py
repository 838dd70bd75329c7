"""python-chess-compatible board over the C ABI host rules (include/chessbot_b200.h, cb_board_*).

Only the subset the reference's path calls (game.py:16,25,44,57-58,85,101,126-128,137,145,154-156;
gameimage.py:37-39,51,54,73,103,121-131,146,168,177-178; puzzles.py:50,54,93,98-99,103).  Importing
this module as `chess` (chessbot_b200/dropin/chess.py) lets the reference's files run unchanged where
python-chess is not installed.
"""
import ctypes as C
import enum

import numpy as np

from ._lib import CB_MAX_MOVES, CB_POS_BYTES, check, lib

WHITE, BLACK = True, False
COLORS = [WHITE, BLACK]
PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = range(1, 7)
PIECE_SYMBOLS = [None, "p", "n", "b", "r", "q", "k"]
SQUARE_NAMES = [f + r for r in "12345678" for f in "abcdefgh"]
STARTING_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"


def parse_square(name):
    return SQUARE_NAMES.index(name)


def square_name(sq):
    return SQUARE_NAMES[sq]


class Termination(enum.Enum):
    CHECKMATE = 1
    STALEMATE = 2
    INSUFFICIENT_MATERIAL = 3
    SEVENTYFIVE_MOVES = 4
    FIVEFOLD_REPETITION = 5
    FIFTY_MOVES = 6
    THREEFOLD_REPETITION = 7


class Outcome:
    def __init__(self, termination, winner):
        self.termination = termination
        self.winner = winner

    def result(self):
        return "1/2-1/2" if self.winner is None else ("1-0" if self.winner else "0-1")


class Move:
    __slots__ = ("from_square", "to_square", "promotion")

    def __init__(self, from_square, to_square, promotion=None):
        self.from_square, self.to_square, self.promotion = from_square, to_square, promotion

    @classmethod
    def from_code(cls, code):
        p = (code >> 12) & 7
        return cls(code & 63, (code >> 6) & 63, p or None)

    @classmethod
    def from_uci(cls, uci):
        if len(uci) not in (4, 5) or uci[:2] not in SQUARE_NAMES or uci[2:4] not in SQUARE_NAMES:
            raise ValueError(f"invalid uci: {uci!r}")
        promo = PIECE_SYMBOLS.index(uci[4]) if len(uci) == 5 else None
        return cls(SQUARE_NAMES.index(uci[:2]), SQUARE_NAMES.index(uci[2:4]), promo)

    def code(self):
        return self.from_square | (self.to_square << 6) | ((self.promotion or 0) << 12)

    def uci(self):
        return SQUARE_NAMES[self.from_square] + SQUARE_NAMES[self.to_square] + (PIECE_SYMBOLS[self.promotion] if self.promotion else "")

    def __eq__(self, o):
        return isinstance(o, Move) and self.code() == o.code()

    def __hash__(self):
        return self.code()

    def __repr__(self):
        return f"Move.from_uci({self.uci()!r})"

    __str__ = uci


def code_to_uci(code):
    return Move.from_code(int(code)).uci()


class SquareSet:
    def __init__(self, mask):
        self.mask = int(mask)

    def __iter__(self):
        m = self.mask
        while m:
            low = m & -m
            yield low.bit_length() - 1
            m ^= low

    def __len__(self):
        return bin(self.mask).count("1")

    def __contains__(self, sq):
        return bool(self.mask >> sq & 1)

    def __int__(self):
        return self.mask


class _LegalMoves:
    def __init__(self, board):
        self._b = board

    def _codes(self):
        buf = (C.c_uint16 * CB_MAX_MOVES)()
        n = lib.cb_board_legal_moves(self._b._h, buf)
        return buf[:n]

    def __iter__(self):
        return iter([Move.from_code(c) for c in self._codes()])

    def __len__(self):
        return len(self._codes())

    count = __len__

    def __bool__(self):
        return len(self) > 0

    def __contains__(self, move):
        return move.code() in self._codes()


class _MoveStack:
    """len() / iteration / indexing of board.move_stack (game.py:16,25; gameimage.py:103,177)."""

    def __init__(self, board):
        self._b = board

    def __len__(self):
        return lib.cb_board_ply(self._b._h)

    def _list(self):
        n = len(self)
        buf = (C.c_uint16 * max(n, 1))()
        lib.cb_board_move_stack(self._b._h, buf, n)
        return [Move.from_code(c) for c in buf[:n]]

    def __iter__(self):
        return iter(self._list())

    def __getitem__(self, i):
        return self._list()[i]


class Board:
    def __init__(self, fen=STARTING_FEN, _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            self._h = lib.cb_board_new(fen.encode() if fen is not None else None)
            if not self._h:
                from ._lib import last_error
                raise ValueError(last_error())

    def __del__(self, _free=lib.cb_board_free):      # bound at definition: module globals are gone at interpreter shutdown
        h, self._h = getattr(self, "_h", None), None
        if h:
            _free(h)

    # --- state
    @property
    def turn(self):
        return bool(lib.cb_board_turn(self._h))

    @property
    def halfmove_clock(self):
        return lib.cb_board_halfmove_clock(self._h)

    @property
    def ep_square(self):
        e = lib.cb_board_ep_square(self._h)
        return None if e < 0 else e

    @property
    def move_stack(self):
        return _MoveStack(self)

    @property
    def legal_moves(self):
        return _LegalMoves(self)

    def pieces_mask(self, piece_type, color):
        return lib.cb_board_pieces_mask(self._h, piece_type, int(bool(color)))

    def pieces(self, piece_type, color):
        return SquareSet(self.pieces_mask(piece_type, color))

    def has_kingside_castling_rights(self, color):
        return bool(lib.cb_board_castling(self._h) & (1 if color else 4))

    def has_queenside_castling_rights(self, color):
        return bool(lib.cb_board_castling(self._h) & (2 if color else 8))

    def is_repetition(self, count=3):
        return bool(lib.cb_board_is_repetition(self._h, count))

    def is_check(self):
        return bool(lib.cb_board_is_check(self._h))

    def fen(self):
        buf = C.create_string_buffer(128)
        lib.cb_board_fen(self._h, buf, 128)
        return buf.value.decode()

    # --- moves
    def push_uci(self, uci):
        if lib.cb_board_push_uci(self._h, uci.encode()) != 0:
            from ._lib import last_error
            raise ValueError(last_error())
        return Move.from_uci(uci)

    def push(self, move):
        return self.push_uci(move.uci())

    def pop(self):
        code = lib.cb_board_pop(self._h)
        if code < 0:
            raise IndexError("pop from empty move stack")
        return Move.from_code(code)

    def copy(self, *, stack=True):
        return Board(_handle=lib.cb_board_copy(self._h))

    def apply_mirror(self):
        lib.cb_board_mirror(self._h)

    def mirror(self):
        b = self.copy()
        b.apply_mirror()
        return b

    # --- game end
    def outcome(self, *, claim_draw=False):
        code = lib.cb_board_outcome(self._h, int(claim_draw))
        if code == 0:
            return None
        term = Termination(code)
        return Outcome(term, (not self.turn) if term is Termination.CHECKMATE else None)

    def is_game_over(self, *, claim_draw=False):
        return lib.cb_board_outcome(self._h, int(claim_draw)) != 0

    def perft(self, depth):
        return lib.cb_board_perft(self._h, depth)

    # --- device export
    def export_record(self):
        """Game record for the device (uint8 [n, 96]): enough positions for the 8 encoder frames and for
        repetition / draw-claim counting back to the last irreversible move."""
        need = lib.cb_board_export_record(self._h, None, 0)
        buf = np.empty((need, CB_POS_BYTES), dtype=np.uint8)
        check(lib.cb_board_export_record(self._h, buf.ctypes.data, need), "export_record")
        return buf

    def __repr__(self):
        return f"Board({self.fen()!r})"


def export_records(boards, flat=False):
    """Game records of many boards in one C call: (list of uint8 [k_i, 96] views, int32 legal-move counts); flat=True returns
    (uint8 [sum k_i, 96], int32 lens, int32 legal-move counts) without cutting the buffer into per-board views."""
    n = len(boards)
    handles = (C.c_void_p * n)(*[b._h for b in boards])
    lens = np.empty(n, dtype=np.int32)
    n_legal = np.empty(n, dtype=np.int32)
    cap = 16 * n
    while True:
        buf = np.empty((cap, CB_POS_BYTES), dtype=np.uint8)
        total = lib.cb_boards_export_records(handles, n, buf.ctypes.data, cap, lens.ctypes.data, n_legal.ctypes.data)
        if total <= cap:
            break
        cap = total
    if flat:
        return buf[:total], lens, n_legal
    ends = np.cumsum(lens)
    return [buf[e - k:e] for e, k in zip(ends.tolist(), lens.tolist())], n_legal


def _piece_type_at(board, sq):
    for pt in range(1, 7):
        if (board.pieces_mask(pt, True) | board.pieces_mask(pt, False)) >> sq & 1:
            return pt
    return None


def _parse_san(board, san):
    """SAN parsing for tests/game.py and Game.make_image_sample (game.py:154-156)."""
    s = san.rstrip("+#!?")
    legal = list(board.legal_moves)
    if s in ("O-O", "0-0", "O-O-O", "0-0-0"):
        to_file = 6 if s in ("O-O", "0-0") else 2
        for m in legal:
            if _piece_type_at(board, m.from_square) == KING and abs((m.from_square & 7) - (m.to_square & 7)) == 2 \
                    and (m.to_square & 7) == to_file:
                return m
        raise ValueError(f"illegal san: {san!r}")
    promo = None
    if "=" in s:
        s, p = s.split("=")
        promo = PIECE_SYMBOLS.index(p.lower())
    piece = PAWN
    if s[0] in "NBRQK":
        piece = PIECE_SYMBOLS.index(s[0].lower())
        s = s[1:]
    s = s.replace("x", "")
    to = SQUARE_NAMES.index(s[-2:])
    cands = [m for m in legal if m.to_square == to and m.promotion == promo
             and _piece_type_at(board, m.from_square) == piece and all(ch in SQUARE_NAMES[m.from_square] for ch in s[:-2])]
    if len(cands) != 1:
        raise ValueError(f"{'ambiguous' if cands else 'illegal'} san: {san!r} in {board.fen()}")
    return cands[0]


def _push_san(self, san):
    m = _parse_san(self, san)
    self.push(m)
    return m


Board.push_san = _push_san
Board.parse_san = lambda self, san: _parse_san(self, san)
Board.piece_type_at = _piece_type_at
