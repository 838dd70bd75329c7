"""ORACLE (test infrastructure, not product code).

A from-scratch, pure-Python restatement of the subset of the third-party
``chess`` module (python-chess, un-vendored and un-pinned in the reference;
≈1.10.x by date) that the reference's hot path calls.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

The reference never implements a chess rule itself, so the behaviours below are
part of its observable results:

* ``legal_moves`` iteration ORDER (feeds ``mcts.py:92-101`` children order, the
  f32 sum order ``mcts.py:96``, noise order ``mcts.py:158-161``);
* ``is_game_over(claim_draw=True)`` / ``outcome(claim_draw=True)`` incl. the
  claim-ahead fifty-move and threefold rules (``game.py:57-58``);
* ``is_repetition(2|3)`` (``gameimage.py:73``), cleaned castling rights
  (``gameimage.py:121-127``), ``apply_mirror`` (``gameimage.py:39``),
  ``copy``/``pop``/``move_stack`` (``gameimage.py:168,177-178``; ``game.py:137``).

Call sites covered: ``game.py:16,25,44,57-58,85,101,126-128,137,145,154-156``,
``gameimage.py:37-39,51,54,73,103,121-131,146,168,177-178``,
``puzzles.py:50,54,93,98-99,103``.

Parity pins: perft known answers (public), the 43-move ordered golden list of
``main.ipynb:109-151``, the reference's own ``tests/game.py`` and
``tests/gameimage.py`` (run against this shim by tests/test_oracle_*.py).
Draw-claim / repetition behaviour is restated from the published python-chess
algorithm: PARITY UNPINNED by the reference's own tests for those rules.
"""
from __future__ import annotations

import enum
from typing import Iterator, List, Optional

WHITE = True
BLACK = False
COLORS = [WHITE, BLACK]
PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = range(1, 7)
PIECE_TYPES = [PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING]
PIECE_SYMBOLS = [None, "p", "n", "b", "r", "q", "k"]
FILE_NAMES = "abcdefgh"
RANK_NAMES = "12345678"
STARTING_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"

SQUARES = list(range(64))
SQUARE_NAMES = [f + r for r in RANK_NAMES for f in FILE_NAMES]
(A1, B1, C1, D1, E1, F1, G1, H1) = range(8)
(A8, B8, C8, D8, E8, F8, G8, H8) = range(56, 64)

BB_ALL = 0xFFFF_FFFF_FFFF_FFFF
BB_SQUARES = [1 << s for s in range(64)]
BB_FILES = [0x0101_0101_0101_0101 << f for f in range(8)]
BB_RANKS = [0xFF << (8 * r) for r in range(8)]
BB_LIGHT_SQUARES = 0x55AA_55AA_55AA_55AA
BB_DARK_SQUARES = 0xAA55_AA55_AA55_AA55
BB_BACKRANKS = BB_RANKS[0] | BB_RANKS[7]


def square_file(sq: int) -> int:
    return sq & 7


def square_rank(sq: int) -> int:
    return sq >> 3


def square(file: int, rank: int) -> int:
    return rank * 8 + file


def square_name(sq: int) -> str:
    return SQUARE_NAMES[sq]


def parse_square(name: str) -> int:
    return SQUARE_NAMES.index(name)


def msb(bb: int) -> int:
    return bb.bit_length() - 1


def lsb(bb: int) -> int:
    return (bb & -bb).bit_length() - 1


def popcount(bb: int) -> int:
    return bin(bb).count("1")


def scan_reversed(bb: int) -> Iterator[int]:
    """Squares of a bitboard from h8 down to a1 (the generation order)."""
    while bb:
        s = bb.bit_length() - 1
        yield s
        bb ^= 1 << s


def scan_forward(bb: int) -> Iterator[int]:
    while bb:
        low = bb & -bb
        yield low.bit_length() - 1
        bb ^= low


def flip_vertical(bb: int) -> int:
    out = 0
    for r in range(8):
        out |= ((bb >> (8 * r)) & 0xFF) << (8 * (7 - r))
    return out


def _step_attacks(sq: int, deltas) -> int:
    bb = 0
    f, r = sq & 7, sq >> 3
    for df, dr in deltas:
        nf, nr = f + df, r + dr
        if 0 <= nf < 8 and 0 <= nr < 8:
            bb |= 1 << (nr * 8 + nf)
    return bb


_KNIGHT_D = [(1, 2), (2, 1), (2, -1), (1, -2), (-1, -2), (-2, -1), (-2, 1), (-1, 2)]
_KING_D = [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)]
BB_KNIGHT_ATTACKS = [_step_attacks(s, _KNIGHT_D) for s in range(64)]
BB_KING_ATTACKS = [_step_attacks(s, _KING_D) for s in range(64)]
# BB_PAWN_ATTACKS[color][sq]: squares a pawn of `color` on sq attacks.
BB_PAWN_ATTACKS = {
    WHITE: [_step_attacks(s, [(-1, 1), (1, 1)]) for s in range(64)],
    BLACK: [_step_attacks(s, [(-1, -1), (1, -1)]) for s in range(64)],
}

_ROOK_DIRS = [(0, 1), (1, 0), (0, -1), (-1, 0)]
_BISHOP_DIRS = [(1, 1), (1, -1), (-1, -1), (-1, 1)]


def _ray_squares(sq: int, df: int, dr: int) -> List[int]:
    out = []
    f, r = (sq & 7) + df, (sq >> 3) + dr
    while 0 <= f < 8 and 0 <= r < 8:
        out.append(r * 8 + f)
        f += df
        r += dr
    return out


_ROOK_RAYS = [[_ray_squares(s, df, dr) for df, dr in _ROOK_DIRS] for s in range(64)]
_BISHOP_RAYS = [[_ray_squares(s, df, dr) for df, dr in _BISHOP_DIRS] for s in range(64)]


def _slide(rays, occupied: int) -> int:
    bb = 0
    for ray in rays:
        for s in ray:
            b = 1 << s
            bb |= b
            if occupied & b:
                break
    return bb


def rook_attacks(sq: int, occupied: int) -> int:
    return _slide(_ROOK_RAYS[sq], occupied)


def bishop_attacks(sq: int, occupied: int) -> int:
    return _slide(_BISHOP_RAYS[sq], occupied)


def _build_lines():
    rays = [[0] * 64 for _ in range(64)]
    betw = [[0] * 64 for _ in range(64)]
    for a in range(64):
        for dirs in (_ROOK_DIRS, _BISHOP_DIRS):
            for df, dr in dirs:
                fwd = _ray_squares(a, df, dr)
                back = _ray_squares(a, -df, -dr)
                line = (1 << a)
                for s in fwd + back:
                    line |= 1 << s
                acc = 0
                for s in fwd:
                    rays[a][s] = line
                    betw[a][s] = acc
                    acc |= 1 << s
    return rays, betw


_BB_RAYS, _BB_BETWEEN = _build_lines()


def ray(a: int, b: int) -> int:
    """Whole line (edge to edge) through a and b, 0 if not aligned."""
    return _BB_RAYS[a][b]


def between(a: int, b: int) -> int:
    """Squares strictly between a and b (0 if not aligned)."""
    return _BB_BETWEEN[a][b]


class Termination(enum.Enum):
    CHECKMATE = enum.auto()
    STALEMATE = enum.auto()
    INSUFFICIENT_MATERIAL = enum.auto()
    SEVENTYFIVE_MOVES = enum.auto()
    FIVEFOLD_REPETITION = enum.auto()
    FIFTY_MOVES = enum.auto()
    THREEFOLD_REPETITION = enum.auto()


class Outcome:
    def __init__(self, termination: Termination, winner: Optional[bool]):
        self.termination = termination
        self.winner = winner

    def result(self) -> str:
        return "1/2-1/2" if self.winner is None else ("1-0" if self.winner else "0-1")

    def __repr__(self):
        return f"Outcome(termination={self.termination!r}, winner={self.winner!r})"


class Move:
    __slots__ = ("from_square", "to_square", "promotion")

    def __init__(self, from_square: int, to_square: int, promotion: Optional[int] = None):
        self.from_square = from_square
        self.to_square = to_square
        self.promotion = promotion

    def uci(self) -> str:
        s = SQUARE_NAMES[self.from_square] + SQUARE_NAMES[self.to_square]
        if self.promotion:
            s += PIECE_SYMBOLS[self.promotion]
        return s

    @classmethod
    def from_uci(cls, uci: str) -> "Move":
        if len(uci) not in (4, 5):
            raise ValueError(f"expected uci string to be of length 4 or 5: {uci!r}")
        try:
            f = SQUARE_NAMES.index(uci[0:2])
            t = SQUARE_NAMES.index(uci[2:4])
            promo = PIECE_SYMBOLS.index(uci[4]) if len(uci) == 5 else None
        except ValueError:
            raise ValueError(f"invalid uci: {uci!r}")
        if f == t:
            raise ValueError(f"invalid uci (use 0000 for null moves): {uci!r}")
        return cls(f, t, promo)

    def __eq__(self, other):
        return (isinstance(other, Move) and self.from_square == other.from_square
                and self.to_square == other.to_square and self.promotion == other.promotion)

    def __hash__(self):
        return hash((self.from_square, self.to_square, self.promotion))

    def __repr__(self):
        return f"Move.from_uci({self.uci()!r})"

    def __str__(self):
        return self.uci()


class Piece:
    def __init__(self, piece_type: int, color: bool):
        self.piece_type = piece_type
        self.color = color

    def symbol(self) -> str:
        s = PIECE_SYMBOLS[self.piece_type]
        return s.upper() if self.color else s

    def __eq__(self, other):
        return isinstance(other, Piece) and (self.piece_type, self.color) == (other.piece_type, other.color)

    def __hash__(self):
        return hash((self.piece_type, self.color))


class SquareSet:
    """Ascending iteration, like python-chess (gameimage.py:51,54 only uses membership)."""

    def __init__(self, mask: int = 0):
        self.mask = mask

    def __iter__(self):
        return scan_forward(self.mask)

    def __len__(self):
        return popcount(self.mask)

    def __contains__(self, sq):
        return bool(self.mask >> sq & 1)

    def __int__(self):
        return self.mask

    def __bool__(self):
        return bool(self.mask)


class LegalMoveGenerator:
    def __init__(self, board: "Board"):
        self.board = board

    def __iter__(self):
        return self.board.generate_legal_moves()

    def __len__(self):
        return sum(1 for _ in self.board.generate_legal_moves())

    def count(self):
        return len(self)

    def __bool__(self):
        return any(True for _ in self.board.generate_legal_moves())

    def __contains__(self, move):
        return self.board.is_legal(move)


class _State:
    __slots__ = ("pawns", "knights", "bishops", "rooks", "queens", "kings", "occ_w", "occ_b",
                 "occupied", "turn", "castling_rights", "ep_square", "halfmove_clock", "fullmove_number")

    def __init__(self, b: "Board"):
        self.pawns, self.knights, self.bishops = b.pawns, b.knights, b.bishops
        self.rooks, self.queens, self.kings = b.rooks, b.queens, b.kings
        self.occ_w, self.occ_b = b.occupied_co[WHITE], b.occupied_co[BLACK]
        self.occupied = b.occupied
        self.turn = b.turn
        self.castling_rights = b.castling_rights
        self.ep_square = b.ep_square
        self.halfmove_clock = b.halfmove_clock
        self.fullmove_number = b.fullmove_number

    def restore(self, b: "Board"):
        b.pawns, b.knights, b.bishops = self.pawns, self.knights, self.bishops
        b.rooks, b.queens, b.kings = self.rooks, self.queens, self.kings
        b.occupied_co = {WHITE: self.occ_w, BLACK: self.occ_b}
        b.occupied = self.occupied
        b.turn = self.turn
        b.castling_rights = self.castling_rights
        b.ep_square = self.ep_square
        b.halfmove_clock = self.halfmove_clock
        b.fullmove_number = self.fullmove_number


class Board:
    def __init__(self, fen: Optional[str] = STARTING_FEN):
        self.move_stack: List[Move] = []
        self._stack: List[_State] = []
        self._clear()
        if fen is not None:
            self.set_fen(fen)

    # ------------------------------------------------------------------ setup
    def _clear(self):
        self.pawns = self.knights = self.bishops = self.rooks = self.queens = self.kings = 0
        self.occupied_co = {WHITE: 0, BLACK: 0}
        self.occupied = 0
        self.turn = WHITE
        self.castling_rights = 0
        self.ep_square = None
        self.halfmove_clock = 0
        self.fullmove_number = 1

    def set_fen(self, fen: str) -> None:
        parts = fen.split()
        if not parts:
            raise ValueError("empty fen")
        board_part = parts[0]
        turn_part = parts[1] if len(parts) > 1 else "w"
        castling_part = parts[2] if len(parts) > 2 else "-"
        ep_part = parts[3] if len(parts) > 3 else "-"
        half_part = parts[4] if len(parts) > 4 else "0"
        full_part = parts[5] if len(parts) > 5 else "1"
        rows = board_part.split("/")
        if len(rows) != 8:
            raise ValueError(f"expected 8 rows in position part of fen: {fen!r}")
        self._clear()
        self.move_stack = []
        self._stack = []
        for i, row in enumerate(rows):
            r = 7 - i
            f = 0
            for ch in row:
                if ch.isdigit():
                    f += int(ch)
                else:
                    pt = PIECE_SYMBOLS.index(ch.lower())
                    self._set_piece_at(r * 8 + f, pt, ch.isupper())
                    f += 1
            if f != 8:
                raise ValueError(f"bad row in fen: {fen!r}")
        if turn_part not in ("w", "b"):
            raise ValueError(f"bad turn in fen: {fen!r}")
        self.turn = turn_part == "w"
        cr = 0
        if castling_part != "-":
            for ch in castling_part:
                if ch == "K":
                    cr |= BB_SQUARES[H1]
                elif ch == "Q":
                    cr |= BB_SQUARES[A1]
                elif ch == "k":
                    cr |= BB_SQUARES[H8]
                elif ch == "q":
                    cr |= BB_SQUARES[A8]
                else:
                    raise ValueError(f"bad castling in fen: {fen!r}")
        self.castling_rights = cr
        self.ep_square = None if ep_part == "-" else SQUARE_NAMES.index(ep_part)
        self.halfmove_clock = int(half_part)
        self.fullmove_number = max(int(full_part), 1)

    def _set_piece_at(self, sq: int, pt: int, color: bool) -> None:
        self._remove_piece_at(sq)
        b = 1 << sq
        if pt == PAWN:
            self.pawns |= b
        elif pt == KNIGHT:
            self.knights |= b
        elif pt == BISHOP:
            self.bishops |= b
        elif pt == ROOK:
            self.rooks |= b
        elif pt == QUEEN:
            self.queens |= b
        else:
            self.kings |= b
        self.occupied |= b
        self.occupied_co[color] |= b

    def _remove_piece_at(self, sq: int) -> Optional[int]:
        pt = self.piece_type_at(sq)
        if pt is None:
            return None
        m = ~(1 << sq) & BB_ALL
        self.pawns &= m
        self.knights &= m
        self.bishops &= m
        self.rooks &= m
        self.queens &= m
        self.kings &= m
        self.occupied &= m
        self.occupied_co[WHITE] &= m
        self.occupied_co[BLACK] &= m
        return pt

    # ---------------------------------------------------------------- queries
    def piece_type_at(self, sq: int) -> Optional[int]:
        b = 1 << sq
        if not self.occupied & b:
            return None
        if self.pawns & b:
            return PAWN
        if self.knights & b:
            return KNIGHT
        if self.bishops & b:
            return BISHOP
        if self.rooks & b:
            return ROOK
        if self.queens & b:
            return QUEEN
        return KING

    def piece_at(self, sq: int) -> Optional[Piece]:
        pt = self.piece_type_at(sq)
        if pt is None:
            return None
        return Piece(pt, bool(self.occupied_co[WHITE] >> sq & 1))

    def color_at(self, sq: int) -> Optional[bool]:
        b = 1 << sq
        if self.occupied_co[WHITE] & b:
            return WHITE
        if self.occupied_co[BLACK] & b:
            return BLACK
        return None

    def pieces_mask(self, piece_type: int, color: bool) -> int:
        bb = (None, self.pawns, self.knights, self.bishops, self.rooks, self.queens, self.kings)[piece_type]
        return bb & self.occupied_co[color]

    def pieces(self, piece_type: int, color: bool) -> SquareSet:
        return SquareSet(self.pieces_mask(piece_type, color))

    def king(self, color: bool) -> Optional[int]:
        k = self.kings & self.occupied_co[color]
        return msb(k) if k else None

    def attacks_mask(self, sq: int) -> int:
        b = 1 << sq
        if b & self.pawns:
            return BB_PAWN_ATTACKS[bool(b & self.occupied_co[WHITE])][sq]
        if b & self.knights:
            return BB_KNIGHT_ATTACKS[sq]
        if b & self.kings:
            return BB_KING_ATTACKS[sq]
        att = 0
        if b & (self.bishops | self.queens):
            att |= bishop_attacks(sq, self.occupied)
        if b & (self.rooks | self.queens):
            att |= rook_attacks(sq, self.occupied)
        return att

    def _attackers_mask(self, color: bool, sq: int, occupied: int) -> int:
        queens_rooks = self.queens | self.rooks
        queens_bishops = self.queens | self.bishops
        att = ((BB_KING_ATTACKS[sq] & self.kings)
               | (BB_KNIGHT_ATTACKS[sq] & self.knights)
               | (rook_attacks(sq, occupied) & queens_rooks)
               | (bishop_attacks(sq, occupied) & queens_bishops)
               | (BB_PAWN_ATTACKS[not color][sq] & self.pawns))
        return att & self.occupied_co[color]

    def attackers_mask(self, color: bool, sq: int) -> int:
        return self._attackers_mask(color, sq, self.occupied)

    def is_attacked_by(self, color: bool, sq: int) -> bool:
        return bool(self.attackers_mask(color, sq))

    def checkers_mask(self) -> int:
        k = self.king(self.turn)
        return 0 if k is None else self.attackers_mask(not self.turn, k)

    def is_check(self) -> bool:
        return bool(self.checkers_mask())

    def pin_mask(self, color: bool, sq: int) -> int:
        k = self.king(color)
        if k is None:
            return BB_ALL
        sq_mask = 1 << sq
        for attacks_fn, sliders in ((rook_attacks, self.rooks | self.queens),
                                    (bishop_attacks, self.bishops | self.queens)):
            line = ray(k, sq)
            # only lines of the right kind
            if not line:
                continue
            if attacks_fn(k, 0) & sq_mask:
                snipers = line & sliders & self.occupied_co[not color] & attacks_fn(k, 0)
                for sniper in scan_reversed(snipers):
                    if between(sniper, k) & (self.occupied | sq_mask) == sq_mask:
                        return line
                break
        return BB_ALL

    # ---------------------------------------------------------- castling etc.
    def clean_castling_rights(self) -> int:
        if self._stack:
            # rights are filtered at every push, so they are already clean
            return self.castling_rights
        castling = self.castling_rights & self.rooks
        white = castling & BB_RANKS[0] & self.occupied_co[WHITE] & (BB_SQUARES[A1] | BB_SQUARES[H1])
        black = castling & BB_RANKS[7] & self.occupied_co[BLACK] & (BB_SQUARES[A8] | BB_SQUARES[H8])
        if not self.occupied_co[WHITE] & self.kings & BB_SQUARES[E1]:
            white = 0
        if not self.occupied_co[BLACK] & self.kings & BB_SQUARES[E8]:
            black = 0
        return white | black

    def has_castling_rights(self, color: bool) -> bool:
        return bool(self.clean_castling_rights() & (BB_RANKS[0] if color else BB_RANKS[7]))

    def has_kingside_castling_rights(self, color: bool) -> bool:
        return bool(self.clean_castling_rights() & BB_SQUARES[H1 if color else H8])

    def has_queenside_castling_rights(self, color: bool) -> bool:
        return bool(self.clean_castling_rights() & BB_SQUARES[A1 if color else A8])

    def is_en_passant(self, move: Move) -> bool:
        return (self.ep_square == move.to_square
                and bool(self.pawns & (1 << move.from_square))
                and abs(move.to_square - move.from_square) in (7, 9)
                and not self.occupied & (1 << move.to_square))

    def is_capture(self, move: Move) -> bool:
        return bool((1 << move.to_square) & self.occupied_co[not self.turn]) or self.is_en_passant(move)

    def is_zeroing(self, move: Move) -> bool:
        touched = (1 << move.from_square) ^ (1 << move.to_square)
        return bool(touched & self.pawns or touched & self.occupied_co[not self.turn])

    def is_castling(self, move: Move) -> bool:
        if self.kings & (1 << move.from_square):
            diff = square_file(move.from_square) - square_file(move.to_square)
            return abs(diff) > 1 or bool(self.rooks & self.occupied_co[self.turn] & (1 << move.to_square))
        return False

    def _reduces_castling_rights(self, move: Move) -> bool:
        cr = self.clean_castling_rights()
        touched = (1 << move.from_square) ^ (1 << move.to_square)
        return bool(touched & cr
                    or (cr & BB_RANKS[0] and touched & self.kings & self.occupied_co[WHITE])
                    or (cr & BB_RANKS[7] and touched & self.kings & self.occupied_co[BLACK]))

    def is_irreversible(self, move: Move) -> bool:
        return self.is_zeroing(move) or self._reduces_castling_rights(move) or self.has_legal_en_passant()

    # ------------------------------------------------------------ generation
    def generate_pseudo_legal_moves(self, from_mask: int = BB_ALL, to_mask: int = BB_ALL) -> Iterator[Move]:
        our = self.occupied_co[self.turn]
        # 1. non-pawn pieces: from-squares descending, targets descending
        for frm in scan_reversed(our & ~self.pawns & from_mask):
            for to in scan_reversed(self.attacks_mask(frm) & ~our & to_mask):
                yield Move(frm, to)
        # 2. castling
        if from_mask & self.kings:
            yield from self.generate_castling_moves(from_mask, to_mask)
        pawns = self.pawns & our & from_mask
        if not pawns:
            return
        # 3. pawn captures (promotions expand q, r, b, n)
        for frm in scan_reversed(pawns):
            targets = BB_PAWN_ATTACKS[self.turn][frm] & self.occupied_co[not self.turn] & to_mask
            for to in scan_reversed(targets):
                if square_rank(to) in (0, 7):
                    for promo in (QUEEN, ROOK, BISHOP, KNIGHT):
                        yield Move(frm, to, promo)
                else:
                    yield Move(frm, to)
        # 4./5. single then double advances, iterated by target square descending
        if self.turn == WHITE:
            single = (pawns << 8) & ~self.occupied & BB_ALL
            double = (single << 8) & ~self.occupied & (BB_RANKS[2] | BB_RANKS[3]) & BB_ALL
        else:
            single = (pawns >> 8) & ~self.occupied
            double = (single >> 8) & ~self.occupied & (BB_RANKS[5] | BB_RANKS[4])
        single &= to_mask
        double &= to_mask
        back = -8 if self.turn == WHITE else 8
        for to in scan_reversed(single):
            frm = to + back
            if square_rank(to) in (0, 7):
                for promo in (QUEEN, ROOK, BISHOP, KNIGHT):
                    yield Move(frm, to, promo)
            else:
                yield Move(frm, to)
        for to in scan_reversed(double):
            yield Move(to + 2 * back, to)
        # 6. en passant
        if self.ep_square is not None:
            yield from self.generate_pseudo_legal_ep(from_mask, to_mask)

    def generate_pseudo_legal_ep(self, from_mask: int = BB_ALL, to_mask: int = BB_ALL) -> Iterator[Move]:
        if self.ep_square is None or not BB_SQUARES[self.ep_square] & to_mask:
            return
        if BB_SQUARES[self.ep_square] & self.occupied:
            return
        capturers = (self.pawns & self.occupied_co[self.turn] & from_mask
                     & BB_PAWN_ATTACKS[not self.turn][self.ep_square]
                     & BB_RANKS[4 if self.turn else 3])
        for c in scan_reversed(capturers):
            yield Move(c, self.ep_square)

    def _attacked_for_king(self, path: int, occupied: int) -> bool:
        return any(self._attackers_mask(not self.turn, sq, occupied) for sq in scan_reversed(path))

    def generate_castling_moves(self, from_mask: int = BB_ALL, to_mask: int = BB_ALL) -> Iterator[Move]:
        backrank = BB_RANKS[0] if self.turn == WHITE else BB_RANKS[7]
        king = self.occupied_co[self.turn] & self.kings & backrank & from_mask
        king &= -king
        if not king:
            return
        ksq = msb(king)
        bb_c = BB_FILES[2] & backrank
        bb_d = BB_FILES[3] & backrank
        bb_f = BB_FILES[5] & backrank
        bb_g = BB_FILES[6] & backrank
        # rook squares descending: h-side (king side) first
        for candidate in scan_reversed(self.clean_castling_rights() & backrank & to_mask):
            rook = BB_SQUARES[candidate]
            a_side = rook < king
            king_to = bb_c if a_side else bb_g
            rook_to = bb_d if a_side else bb_f
            king_path = between(ksq, msb(king_to))
            rook_path = between(candidate, msb(rook_to))
            if not ((self.occupied ^ king ^ rook) & (king_path | rook_path | king_to | rook_to)
                    or self._attacked_for_king(king_path | king, self.occupied ^ king)
                    or self._attacked_for_king(king_to, self.occupied ^ king ^ rook ^ rook_to)):
                # standard chess: encoded king-from -> king-to (e1g1 / e1c1)
                yield Move(ksq, msb(king_to))

    def _slider_blockers(self, king: int) -> int:
        rooks_and_queens = self.rooks | self.queens
        bishops_and_queens = self.bishops | self.queens
        snipers = ((rook_attacks(king, 0) & rooks_and_queens)
                   | (bishop_attacks(king, 0) & bishops_and_queens))
        blockers = 0
        for sniper in scan_reversed(snipers & self.occupied_co[not self.turn]):
            b = between(king, sniper) & self.occupied
            # exactly one piece in between
            if b and BB_SQUARES[msb(b)] == b:
                blockers |= b
        return blockers & self.occupied_co[self.turn]

    def _ep_skewered(self, king: int, capturer: int) -> bool:
        # the capturer and the captured pawn both leave the king's rank/diagonal
        last_double = self.ep_square + (-8 if self.turn == WHITE else 8)
        occupancy = (self.occupied & ~BB_SQUARES[last_double] & ~BB_SQUARES[capturer]) | BB_SQUARES[self.ep_square]
        horizontal = self.occupied_co[not self.turn] & (self.rooks | self.queens)
        if rook_attacks(king, occupancy) & BB_RANKS[square_rank(king)] & horizontal:
            return True
        diagonal = self.occupied_co[not self.turn] & (self.bishops | self.queens)
        if bishop_attacks(king, occupancy) & diagonal:
            return True
        return False

    def _is_safe(self, king: int, blockers: int, move: Move) -> bool:
        if move.from_square == king:
            if self.is_castling(move):
                return True
            return not self.is_attacked_by(not self.turn, move.to_square)
        if self.is_en_passant(move):
            return bool(self.pin_mask(self.turn, move.from_square) & BB_SQUARES[move.to_square]
                        and not self._ep_skewered(king, move.from_square))
        return bool(not blockers & BB_SQUARES[move.from_square]
                    or ray(move.from_square, move.to_square) & BB_SQUARES[king])

    def _generate_evasions(self, king: int, checkers: int, from_mask: int = BB_ALL, to_mask: int = BB_ALL):
        sliders = checkers & (self.bishops | self.rooks | self.queens)
        attacked = 0
        for checker in scan_reversed(sliders):
            attacked |= ray(king, checker) & ~BB_SQUARES[checker]
        if BB_SQUARES[king] & from_mask:
            for to in scan_reversed(BB_KING_ATTACKS[king] & ~self.occupied_co[self.turn] & ~attacked & to_mask):
                yield Move(king, to)
        checker = msb(checkers)
        if BB_SQUARES[checker] == checkers:
            # single checker: capture it or block
            target = between(king, checker) | checkers
            yield from self.generate_pseudo_legal_moves(~self.kings & from_mask, target & to_mask)
            # capture the checking pawn en passant (avoid duplicates)
            if self.ep_square is not None and not BB_SQUARES[self.ep_square] & target:
                last_double = self.ep_square + (-8 if self.turn == WHITE else 8)
                if last_double == checker:
                    yield from self.generate_pseudo_legal_ep(from_mask, to_mask)

    def generate_legal_moves(self, from_mask: int = BB_ALL, to_mask: int = BB_ALL) -> Iterator[Move]:
        king_mask = self.kings & self.occupied_co[self.turn]
        if king_mask:
            king = msb(king_mask)
            blockers = self._slider_blockers(king)
            checkers = self.attackers_mask(not self.turn, king)
            gen = (self._generate_evasions(king, checkers, from_mask, to_mask) if checkers
                   else self.generate_pseudo_legal_moves(from_mask, to_mask))
            for move in gen:
                if self._is_safe(king, blockers, move):
                    yield move
        else:
            yield from self.generate_pseudo_legal_moves(from_mask, to_mask)

    def generate_legal_ep(self) -> Iterator[Move]:
        if self.ep_square is None:
            return
        legal = list(self.generate_legal_moves(BB_ALL, BB_SQUARES[self.ep_square]))
        for m in legal:
            if self.is_en_passant(m):
                yield m

    def has_legal_en_passant(self) -> bool:
        return self.ep_square is not None and any(True for _ in self.generate_legal_ep())

    @property
    def legal_moves(self) -> LegalMoveGenerator:
        return LegalMoveGenerator(self)

    def is_legal(self, move: Move) -> bool:
        return any(m == move for m in self.generate_legal_moves())

    # ----------------------------------------------------------- make/unmake
    def push(self, move: Move) -> None:
        state = _State(self)
        clean_rights = self.clean_castling_rights()  # evaluated on the pre-move board
        self._stack.append(state)
        self.move_stack.append(move)
        ep_square = self.ep_square
        self.ep_square = None
        # clocks
        self.halfmove_clock = 0 if self.is_zeroing(move) else self.halfmove_clock + 1
        if self.turn == BLACK:
            self.fullmove_number += 1
        from_bb = 1 << move.from_square
        to_bb = 1 << move.to_square
        piece_type = self._remove_piece_at(move.from_square)
        assert piece_type is not None, f"push() expects a piece on {SQUARE_NAMES[move.from_square]}"
        capture_square = move.to_square
        captured = self.piece_type_at(capture_square)
        self.castling_rights = clean_rights & ~to_bb & ~from_bb
        if piece_type == KING:
            self.castling_rights &= ~(BB_RANKS[0] if self.turn == WHITE else BB_RANKS[7])
        castling = False
        if piece_type == KING:
            diff = square_file(move.to_square) - square_file(move.from_square)
            castling = abs(diff) > 1
        if piece_type == PAWN:
            diff = move.to_square - move.from_square
            if diff == 16 and square_rank(move.from_square) == 1:
                self.ep_square = move.from_square + 8
            elif diff == -16 and square_rank(move.from_square) == 6:
                self.ep_square = move.from_square - 8
            elif move.to_square == ep_square and abs(diff) in (7, 9) and not captured:
                down = -8 if self.turn == WHITE else 8
                capture_square = ep_square + down
                captured = self._remove_piece_at(capture_square)
        if move.promotion:
            piece_type = move.promotion
        if castling:
            a_side = square_file(move.to_square) < square_file(move.from_square)
            rank0 = 0 if self.turn == WHITE else 56
            rook_from = rank0 + (0 if a_side else 7)
            rook_to = rank0 + (3 if a_side else 5)
            self._remove_piece_at(rook_from)
            self._set_piece_at(move.to_square, KING, self.turn)
            self._set_piece_at(rook_to, ROOK, self.turn)
        else:
            self._set_piece_at(move.to_square, piece_type, self.turn)
        self.turn = not self.turn

    def pop(self) -> Move:
        move = self.move_stack.pop()
        self._stack.pop().restore(self)
        return move

    def peek(self) -> Move:
        return self.move_stack[-1]

    def parse_uci(self, uci: str) -> Move:
        move = Move.from_uci(uci)
        if not self.is_legal(move):
            raise ValueError(f"illegal uci: {uci!r} in {self.fen()}")
        return move

    def push_uci(self, uci: str) -> Move:
        move = self.parse_uci(uci)
        self.push(move)
        return move

    # SAN (tests/game.py, game.py:154-156 only)
    def parse_san(self, san: str) -> Move:
        s = san.rstrip("+#!?")
        legal = list(self.generate_legal_moves())
        if s in ("O-O", "0-0", "O-O-O", "0-0-0"):
            kingside = s in ("O-O", "0-0")
            for m in legal:
                if self.is_castling(m) and (square_file(m.to_square) == 6) == kingside:
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
        dis = s[:-2]
        cands = []
        for m in legal:
            if m.to_square != to or m.promotion != promo or self.piece_type_at(m.from_square) != piece:
                continue
            name = SQUARE_NAMES[m.from_square]
            if all(ch in name for ch in dis):
                cands.append(m)
        if len(cands) != 1:
            raise ValueError(f"{'ambiguous' if cands else 'illegal'} san: {san!r} in {self.fen()}")
        return cands[0]

    def push_san(self, san: str) -> Move:
        move = self.parse_san(san)
        self.push(move)
        return move

    # ------------------------------------------------------------- game end
    def is_checkmate(self) -> bool:
        return self.is_check() and not any(True for _ in self.generate_legal_moves())

    def is_stalemate(self) -> bool:
        return not self.is_check() and not any(True for _ in self.generate_legal_moves())

    def has_insufficient_material(self, color: bool) -> bool:
        ours = self.occupied_co[color]
        if ours & (self.pawns | self.rooks | self.queens):
            return False
        if ours & self.knights:
            return popcount(ours) <= 2 and not (self.occupied_co[not color] & ~self.kings & ~self.queens)
        if ours & self.bishops:
            same_color = (not self.bishops & BB_DARK_SQUARES) or (not self.bishops & BB_LIGHT_SQUARES)
            return bool(same_color and not self.pawns and not self.knights)
        return True

    def is_insufficient_material(self) -> bool:
        return all(self.has_insufficient_material(c) for c in COLORS)

    def _is_halfmoves(self, n: int) -> bool:
        return self.halfmove_clock >= n and any(True for _ in self.generate_legal_moves())

    def is_seventyfive_moves(self) -> bool:
        return self._is_halfmoves(150)

    def is_fifty_moves(self) -> bool:
        return self._is_halfmoves(100)

    def can_claim_fifty_moves(self) -> bool:
        if self.is_fifty_moves():
            return True
        if self.halfmove_clock >= 99:
            for move in list(self.generate_legal_moves()):
                if not self.is_zeroing(move):
                    self.push(move)
                    try:
                        if self.is_fifty_moves():
                            return True
                    finally:
                        self.pop()
        return False

    def _transposition_key(self):
        return (self.pawns, self.knights, self.bishops, self.rooks, self.queens, self.kings,
                self.occupied_co[WHITE], self.occupied_co[BLACK], self.turn,
                self.clean_castling_rights(),
                self.ep_square if self.has_legal_en_passant() else None)

    def is_repetition(self, count: int = 3) -> bool:
        # cheap filter on occupancy only
        maybe = 1
        for state in reversed(self._stack):
            if state.occupied == self.occupied:
                maybe += 1
                if maybe >= count:
                    break
        if maybe < count:
            return False
        key = self._transposition_key()
        switchyard = []
        try:
            while True:
                if count <= 1:
                    return True
                if len(self.move_stack) < count - 1:
                    break
                move = self.pop()
                switchyard.append(move)
                if self.is_irreversible(move):
                    break
                if self._transposition_key() == key:
                    count -= 1
        finally:
            while switchyard:
                self.push(switchyard.pop())
        return False

    def is_fivefold_repetition(self) -> bool:
        return self.is_repetition(5)

    def can_claim_threefold_repetition(self) -> bool:
        key = self._transposition_key()
        counts = {key: 1}
        switchyard = []
        while self.move_stack:
            move = self.pop()
            switchyard.append(move)
            if self.is_irreversible(move):
                break
            k = self._transposition_key()
            counts[k] = counts.get(k, 0) + 1
        while switchyard:
            self.push(switchyard.pop())
        if counts[key] >= 3:
            return True
        # claim-ahead: some legal move reaches a position seen twice already
        for move in list(self.generate_legal_moves()):
            self.push(move)
            try:
                if counts.get(self._transposition_key(), 0) >= 2:
                    return True
            finally:
                self.pop()
        return False

    def can_claim_draw(self) -> bool:
        return self.can_claim_fifty_moves() or self.can_claim_threefold_repetition()

    def outcome(self, *, claim_draw: bool = False) -> Optional[Outcome]:
        if self.is_checkmate():
            return Outcome(Termination.CHECKMATE, not self.turn)
        if self.is_insufficient_material():
            return Outcome(Termination.INSUFFICIENT_MATERIAL, None)
        if not any(True for _ in self.generate_legal_moves()):
            return Outcome(Termination.STALEMATE, None)
        if self.is_seventyfive_moves():
            return Outcome(Termination.SEVENTYFIVE_MOVES, None)
        if self.is_fivefold_repetition():
            return Outcome(Termination.FIVEFOLD_REPETITION, None)
        if claim_draw:
            if self.can_claim_fifty_moves():
                return Outcome(Termination.FIFTY_MOVES, None)
            if self.can_claim_threefold_repetition():
                return Outcome(Termination.THREEFOLD_REPETITION, None)
        return None

    def is_game_over(self, *, claim_draw: bool = False) -> bool:
        return self.outcome(claim_draw=claim_draw) is not None

    def result(self, *, claim_draw: bool = False) -> str:
        o = self.outcome(claim_draw=claim_draw)
        return o.result() if o else "*"

    # ------------------------------------------------------------ transforms
    def copy(self, *, stack=True) -> "Board":
        b = Board(None)
        b.pawns, b.knights, b.bishops = self.pawns, self.knights, self.bishops
        b.rooks, b.queens, b.kings = self.rooks, self.queens, self.kings
        b.occupied_co = dict(self.occupied_co)
        b.occupied = self.occupied
        b.turn = self.turn
        b.castling_rights = self.castling_rights
        b.ep_square = self.ep_square
        b.halfmove_clock = self.halfmove_clock
        b.fullmove_number = self.fullmove_number
        if stack:
            n = len(self.move_stack) if stack is True else int(stack)
            if n:
                b.move_stack = list(self.move_stack[-n:])
                b._stack = list(self._stack[-n:])  # _State objects are immutable snapshots
        return b

    def apply_mirror(self) -> None:
        self.pawns = flip_vertical(self.pawns)
        self.knights = flip_vertical(self.knights)
        self.bishops = flip_vertical(self.bishops)
        self.rooks = flip_vertical(self.rooks)
        self.queens = flip_vertical(self.queens)
        self.kings = flip_vertical(self.kings)
        w, b = flip_vertical(self.occupied_co[WHITE]), flip_vertical(self.occupied_co[BLACK])
        self.occupied_co = {WHITE: b, BLACK: w}
        self.occupied = w | b
        self.castling_rights = flip_vertical(self.castling_rights)
        if self.ep_square is not None:
            self.ep_square ^= 56
        self.turn = not self.turn
        self.move_stack = []
        self._stack = []

    def mirror(self) -> "Board":
        b = self.copy(stack=False)
        b.apply_mirror()
        return b

    # ------------------------------------------------------------- printing
    def board_fen(self) -> str:
        rows = []
        for r in range(7, -1, -1):
            row, empty = "", 0
            for f in range(8):
                p = self.piece_at(r * 8 + f)
                if p is None:
                    empty += 1
                else:
                    if empty:
                        row += str(empty)
                        empty = 0
                    row += p.symbol()
            if empty:
                row += str(empty)
            rows.append(row)
        return "/".join(rows)

    def castling_xfen(self) -> str:
        cr = self.clean_castling_rights()
        s = ""
        if cr & BB_SQUARES[H1]:
            s += "K"
        if cr & BB_SQUARES[A1]:
            s += "Q"
        if cr & BB_SQUARES[H8]:
            s += "k"
        if cr & BB_SQUARES[A8]:
            s += "q"
        return s or "-"

    def fen(self) -> str:
        ep = SQUARE_NAMES[self.ep_square] if self.has_legal_en_passant() else "-"
        return " ".join([self.board_fen(), "w" if self.turn else "b", self.castling_xfen(), ep,
                         str(self.halfmove_clock), str(self.fullmove_number)])

    def __repr__(self):
        return f"Board({self.fen()!r})"

    def __str__(self):
        out = []
        for r in range(7, -1, -1):
            out.append(" ".join((self.piece_at(r * 8 + f).symbol() if self.piece_at(r * 8 + f) else ".")
                                for f in range(8)))
        return "\n".join(out)


def perft(board: Board, depth: int) -> int:
    if depth == 0:
        return 1
    moves = list(board.generate_legal_moves())
    if depth == 1:
        return len(moves)
    n = 0
    for m in moves:
        board.push(m)
        n += perft(board, depth - 1)
        board.pop()
    return n
