// chessbot_b200 — chess rules for the search hot path, one source for host and device.
//
// The reference delegates every rule to python-chess (game.py:16,57-58,85,101,137;
// gameimage.py:37-39,51,73,121-131).  This is a from-scratch bitboard engine that reproduces
// the behaviours its results depend on:
//   * legal-move ORDER: non-pawn pieces (from-squares descending, targets descending), castling
//     (h-side first), pawn captures (promotions q,r,b,n), single pushes, double pushes, en passant;
//     in check: king steps first, then blocks/captures of a single checker            (mcts.py:92)
//   * outcome(claim_draw=True): checkmate, insufficient material, stalemate, 75-move, fivefold,
//     fifty-move claim (incl. claim-ahead at clock 99), threefold claim (incl. claim-ahead) (game.py:57-58)
//   * transposition key with cleaned castling rights and ep square only when an ep capture is legal,
//     repetition counted back to the last irreversible move                            (gameimage.py:73)
// Everything is register/bit arithmetic: no lookup tables, so the same code runs in a CUDA warp
// lane (mcts.cu) and on the host (host_api.cpp) without constant-memory setup.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_HDN __host__ __device__ inline
// Large rule functions stay real calls on the device: fully inlined, move generation alone was ~20,000 SASS
// instructions per call site (0.3 MB), far beyond the instruction cache of an SM running one warp per tree.
#define CB_HDC __host__ __device__ __noinline__ inline
#else
#define CB_HD inline
#define CB_HDN inline
#define CB_HDC inline
#endif

namespace cb {

typedef uint64_t u64;

enum { PAWN = 0, KNIGHT = 1, BISHOP = 2, ROOK = 3, QUEEN = 4, KING = 5 };  // python-chess type - 1
enum { BLACK = 0, WHITE = 1 };
enum { CR_WK = 1, CR_WQ = 2, CR_BK = 4, CR_BQ = 8 };
// outcome codes; names follow python-chess Termination
enum Outcome {
    OUT_NONE = 0, OUT_CHECKMATE = 1, OUT_STALEMATE = 2, OUT_INSUFFICIENT_MATERIAL = 3,
    OUT_SEVENTYFIVE_MOVES = 4, OUT_FIVEFOLD_REPETITION = 5, OUT_FIFTY_MOVES = 6,
    OUT_THREEFOLD_REPETITION = 7, OUT_MAX_LENGTH = 8
};

constexpr int MAX_MOVES = 224;  // 218 is the known maximum of legal moves in a position

typedef uint16_t Move;  // from | to << 6 | promo << 12 (promo: 0 none, python-chess type 2..5 = n,b,r,q)
CB_HD Move make_move(int from, int to, int promo = 0) { return (Move)(from | (to << 6) | (promo << 12)); }
CB_HD int move_from(Move m) { return m & 63; }
CB_HD int move_to(Move m) { return (m >> 6) & 63; }
CB_HD int move_promo(Move m) { return (m >> 12) & 7; }

// One position of a game record: 96 bytes, the unit the encoder and the tree read.
struct Pos {
    u64 bb[6];        // pawns, knights, bishops, rooks, queens, kings (both colours)
    u64 occ[2];       // [BLACK], [WHITE]
    u64 hash;         // 64-bit mix of the transposition key (quick reject; equality is checked in full)
    uint16_t halfmove;  // halfmove clock
    uint16_t ply;       // len(move_stack) of the python-chess board (0 for a FEN root)
    uint8_t turn;       // WHITE / BLACK
    uint8_t castling;   // cleaned rights, CR_* bits
    int8_t ep;          // raw en-passant square or -1
    int8_t ep_key;      // ep square if an en-passant capture is legal, else -1 (part of the key)
    uint8_t irrev_in;   // the move INTO this position was irreversible (or: start of the record)
    uint8_t occur;      // occurrences of this key since the last irreversible move, incl. this one (<=255)
    Move last_move;     // move that led here (0 for the start of the record)
    uint32_t pad_[3];
};
static_assert(sizeof(Pos) == 96, "Pos must be 96 bytes");

// ---------------------------------------------------------------------------------- bit helpers
CB_HD int popcnt(u64 x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
CB_HD int msb(u64 x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)x);
#else
    return 63 - __builtin_clzll(x);
#endif
}
CB_HD int lsb(u64 x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
CB_HD u64 bitrev(u64 x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}
CB_HD u64 bit(int sq) { return 1ULL << sq; }

constexpr u64 FILE_A = 0x0101010101010101ULL;
constexpr u64 FILE_H = 0x8080808080808080ULL;
constexpr u64 RANK_1 = 0xFFULL;
constexpr u64 RANK_8 = 0xFF00000000000000ULL;
constexpr u64 DIAG_MAIN = 0x8040201008040201ULL;  // a1-h8
constexpr u64 DIAG_ANTI = 0x0102040810204080ULL;  // h1-a8
constexpr u64 DARK_SQUARES = 0xAA55AA55AA55AA55ULL;
constexpr u64 LIGHT_SQUARES = 0x55AA55AA55AA55AAULL;

CB_HD u64 rank_mask(int sq) { return RANK_1 << (sq & 56); }
CB_HD u64 file_mask(int sq) { return FILE_A << (sq & 7); }
CB_HD u64 diag_mask(int sq) {
    int d = (sq & 7) - (sq >> 3);
    return d >= 0 ? DIAG_MAIN >> (8 * d) : DIAG_MAIN << (8 * -d);
}
CB_HD u64 anti_mask(int sq) {
    int d = (sq & 7) + (sq >> 3) - 7;
    return d >= 0 ? DIAG_ANTI << (8 * d) : DIAG_ANTI >> (8 * -d);
}
// sliding attacks along one line (hyperbola quintessence with a full bit reversal)
CB_HD u64 line_attacks(int sq, u64 occ, u64 mask) {
    u64 s = bit(sq);
    u64 o = occ & mask;
    u64 fwd = o - 2 * s;
    u64 rev = bitrev(bitrev(o) - 2 * bitrev(s));
    return (fwd ^ rev) & mask;
}
CB_HD u64 rook_attacks(int sq, u64 occ) { return line_attacks(sq, occ, rank_mask(sq)) | line_attacks(sq, occ, file_mask(sq)); }
CB_HD u64 bishop_attacks(int sq, u64 occ) { return line_attacks(sq, occ, diag_mask(sq)) | line_attacks(sq, occ, anti_mask(sq)); }
CB_HD u64 knight_attacks(int sq) {
    u64 b = bit(sq);
    u64 l1 = (b >> 1) & ~FILE_H, l2 = (b >> 2) & 0x3F3F3F3F3F3F3F3FULL;
    u64 r1 = (b << 1) & ~FILE_A, r2 = (b << 2) & 0xFCFCFCFCFCFCFCFCULL;
    u64 h1 = l1 | r1, h2 = l2 | r2;
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
CB_HD u64 king_attacks(int sq) {
    u64 b = bit(sq);
    u64 h = ((b >> 1) & ~FILE_H) | ((b << 1) & ~FILE_A);
    u64 row = h | b;
    return h | (row << 8) | (row >> 8);
}
// squares a pawn of `color` standing on sq attacks
CB_HD u64 pawn_attacks(int color, int sq) {
    u64 b = bit(sq);
    u64 h = ((b >> 1) & ~FILE_H) | ((b << 1) & ~FILE_A);
    return color == WHITE ? h << 8 : h >> 8;
}
// whole line through a and b (edge to edge), 0 if not aligned
CB_HD u64 line_through(int a, int b) {
    if (a == b) return 0;
    u64 bb = bit(b);
    u64 m = rank_mask(a);
    if (m & bb) return m;
    m = file_mask(a);
    if (m & bb) return m;
    m = diag_mask(a);
    if (m & bb) return m;
    m = anti_mask(a);
    if (m & bb) return m;
    return 0;
}
// squares strictly between a and b, 0 if not aligned
CB_HD u64 between(int a, int b) {
    u64 l = line_through(a, b);
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return l & (bit(hi) - 1) & ~(bit(lo) | (bit(lo) - 1));
}

// ---------------------------------------------------------------------------------- queries
CB_HD u64 occupied(const Pos& p) { return p.occ[0] | p.occ[1]; }
CB_HD int piece_type_at(const Pos& p, int sq) {  // -1 if empty
    u64 b = bit(sq);
    if (!(occupied(p) & b)) return -1;
    for (int t = 0; t < 5; ++t)
        if (p.bb[t] & b) return t;
    return KING;
}
CB_HDC u64 attackers_mask(const Pos& p, int color, int sq, u64 occ) {
    u64 qr = p.bb[QUEEN] | p.bb[ROOK], qb = p.bb[QUEEN] | p.bb[BISHOP];
    u64 a = (king_attacks(sq) & p.bb[KING]) | (knight_attacks(sq) & p.bb[KNIGHT]) |
            (rook_attacks(sq, occ) & qr) | (bishop_attacks(sq, occ) & qb) |
            (pawn_attacks(color ^ 1, sq) & p.bb[PAWN]);
    return a & p.occ[color];
}
CB_HD int king_square(const Pos& p, int color) {
    u64 k = p.bb[KING] & p.occ[color];
    return k ? msb(k) : -1;
}
CB_HD u64 checkers_mask(const Pos& p) {
    int k = king_square(p, p.turn);
    return k < 0 ? 0 : attackers_mask(p, p.turn ^ 1, k, occupied(p));
}
CB_HD u64 piece_attacks(const Pos& p, int sq) {
    u64 b = bit(sq);
    u64 occ = occupied(p);
    if (b & p.bb[PAWN]) return pawn_attacks((p.occ[WHITE] & b) ? WHITE : BLACK, sq);
    if (b & p.bb[KNIGHT]) return knight_attacks(sq);
    if (b & p.bb[KING]) return king_attacks(sq);
    u64 a = 0;
    if (b & (p.bb[BISHOP] | p.bb[QUEEN])) a |= bishop_attacks(sq, occ);
    if (b & (p.bb[ROOK] | p.bb[QUEEN])) a |= rook_attacks(sq, occ);
    return a;
}
CB_HD u64 castling_rook_squares(const Pos& p) {  // rights as a rook-square bitboard (python-chess form)
    u64 r = 0;
    if (p.castling & CR_WK) r |= bit(7);
    if (p.castling & CR_WQ) r |= bit(0);
    if (p.castling & CR_BK) r |= bit(63);
    if (p.castling & CR_BQ) r |= bit(56);
    return r;
}
CB_HD uint8_t clean_castling(const Pos& p, uint8_t rights) {
    uint8_t c = 0;
    bool wk = (p.bb[KING] & p.occ[WHITE] & bit(4)) != 0, bk = (p.bb[KING] & p.occ[BLACK] & bit(60)) != 0;
    u64 wr = p.bb[ROOK] & p.occ[WHITE], br = p.bb[ROOK] & p.occ[BLACK];
    if ((rights & CR_WK) && wk && (wr & bit(7))) c |= CR_WK;
    if ((rights & CR_WQ) && wk && (wr & bit(0))) c |= CR_WQ;
    if ((rights & CR_BK) && bk && (br & bit(63))) c |= CR_BK;
    if ((rights & CR_BQ) && bk && (br & bit(56))) c |= CR_BQ;
    return c;
}
CB_HD bool is_en_passant(const Pos& p, Move m) {
    int f = move_from(m), t = move_to(m), d = t - f;
    if (d < 0) d = -d;
    return p.ep == t && (p.bb[PAWN] & bit(f)) && (d == 7 || d == 9) && !(occupied(p) & bit(t));
}
CB_HD bool is_zeroing(const Pos& p, Move m) {  // pawn move or capture
    u64 touched = bit(move_from(m)) ^ bit(move_to(m));
    return (touched & p.bb[PAWN]) || (touched & p.occ[p.turn ^ 1]);
}
CB_HD bool is_castling_move(const Pos& p, Move m) {
    if (!(p.bb[KING] & bit(move_from(m)))) return false;
    int d = (move_from(m) & 7) - (move_to(m) & 7);
    return d > 1 || d < -1;
}

// ---------------------------------------------------------------------------------- generation
struct MoveList {
    Move m[MAX_MOVES];
    int n;
};

// blockers: own pieces that alone stand between the king and an enemy slider
CB_HDC u64 slider_blockers(const Pos& p, int king) {
    u64 snipers = ((rook_attacks(king, 0) & (p.bb[ROOK] | p.bb[QUEEN])) |
                   (bishop_attacks(king, 0) & (p.bb[BISHOP] | p.bb[QUEEN]))) & p.occ[p.turn ^ 1];
    u64 occ = occupied(p), blockers = 0;
    while (snipers) {
        int s = msb(snipers);
        snipers ^= bit(s);
        u64 b = between(king, s) & occ;
        if (b && !(b & (b - 1))) blockers |= b;
    }
    return blockers & p.occ[p.turn];
}
CB_HD bool ep_skewered(const Pos& p, int king, int capturer) {
    int last_double = p.ep + (p.turn == WHITE ? -8 : 8);
    u64 occ = (occupied(p) & ~bit(last_double) & ~bit(capturer)) | bit(p.ep);
    u64 them = p.occ[p.turn ^ 1];
    if (rook_attacks(king, occ) & rank_mask(king) & them & (p.bb[ROOK] | p.bb[QUEEN])) return true;
    if (bishop_attacks(king, occ) & them & (p.bb[BISHOP] | p.bb[QUEEN])) return true;
    return false;
}
// line on which the piece on sq is pinned to its king, or all ones
CB_HDC u64 pin_mask(const Pos& p, int color, int sq) {
    int king = king_square(p, color);
    if (king < 0) return ~0ULL;
    u64 line = line_through(king, sq);
    if (!line) return ~0ULL;
    bool straight = (line == rank_mask(king)) || (line == file_mask(king));
    u64 sliders = straight ? (p.bb[ROOK] | p.bb[QUEEN]) : (p.bb[BISHOP] | p.bb[QUEEN]);
    u64 snipers = line & sliders & p.occ[color ^ 1];
    u64 occ = occupied(p) | bit(sq);
    while (snipers) {
        int s = msb(snipers);
        snipers ^= bit(s);
        if ((between(s, king) & occ) == bit(sq)) return line;
    }
    return ~0ULL;
}
CB_HDC bool is_safe(const Pos& p, int king, u64 blockers, Move m) {
    int f = move_from(m), t = move_to(m);
    if (f == king) {
        if (is_castling_move(p, m)) return true;
        return attackers_mask(p, p.turn ^ 1, t, occupied(p)) == 0;
    }
    if (is_en_passant(p, m)) return (pin_mask(p, p.turn, f) & bit(t)) && !ep_skewered(p, king, f);
    return !(blockers & bit(f)) || (line_through(f, t) & bit(king));
}

// Emits pseudo-legal moves in python-chess order, filtered by `is_safe` when king >= 0.
struct Gen {
    const Pos& p;
    MoveList& out;
    int king;      // -1: no legality filter
    u64 blockers;
    CB_HD Gen(const Pos& p_, MoveList& o, int k, u64 b) : p(p_), out(o), king(k), blockers(b) {}
    CB_HD void emit(Move m) {
        if (king < 0 || is_safe(p, king, blockers, m)) out.m[out.n++] = m;
    }
    CB_HD void promos_or_plain(int f, int t) {
        if ((t >> 3) == 0 || (t >> 3) == 7) {
            emit(make_move(f, t, 5)); emit(make_move(f, t, 4)); emit(make_move(f, t, 3)); emit(make_move(f, t, 2));
        } else emit(make_move(f, t));
    }
    CB_HD void ep_moves(u64 from_mask, u64 to_mask) {
        if (p.ep < 0 || !(bit(p.ep) & to_mask) || (bit(p.ep) & occupied(p))) return;
        u64 capturers = p.bb[PAWN] & p.occ[p.turn] & from_mask & pawn_attacks(p.turn ^ 1, p.ep) &
                        (p.turn == WHITE ? (RANK_1 << 32) : (RANK_1 << 24));
        while (capturers) {
            int c = msb(capturers);
            capturers ^= bit(c);
            emit(make_move(c, p.ep));
        }
    }
    CB_HD bool attacked_for_king(u64 path, u64 occ) {
        while (path) {
            int s = msb(path);
            path ^= bit(s);
            if (attackers_mask(p, p.turn ^ 1, s, occ)) return true;
        }
        return false;
    }
    CB_HD void castling_moves(u64 from_mask, u64 to_mask) {
        u64 backrank = p.turn == WHITE ? RANK_1 : RANK_8;
        u64 king_bb = p.occ[p.turn] & p.bb[KING] & backrank & from_mask;
        king_bb &= (~king_bb + 1);
        if (!king_bb) return;
        int ksq = msb(king_bb);
        u64 occ = occupied(p);
        u64 cands = castling_rook_squares(p) & backrank & to_mask;
        while (cands) {  // rook squares descending: h-side first
            int c = msb(cands);
            cands ^= bit(c);
            u64 rook = bit(c);
            bool a_side = rook < king_bb;
            int base = p.turn == WHITE ? 0 : 56;
            int king_to = base + (a_side ? 2 : 6), rook_to = base + (a_side ? 3 : 5);
            u64 king_path = between(ksq, king_to), rook_path = between(c, rook_to);
            if (((occ ^ king_bb ^ rook) & (king_path | rook_path | bit(king_to) | bit(rook_to))) ||
                attacked_for_king(king_path | king_bb, occ ^ king_bb) ||
                attacked_for_king(bit(king_to), occ ^ king_bb ^ rook ^ bit(rook_to)))
                continue;
            emit(make_move(ksq, king_to));
        }
    }
    CB_HD void pseudo_legal(u64 from_mask, u64 to_mask) {
        u64 our = p.occ[p.turn], them = p.occ[p.turn ^ 1], occ = our | them;
        u64 pieces = our & ~p.bb[PAWN] & from_mask;
        while (pieces) {
            int f = msb(pieces);
            pieces ^= bit(f);
            u64 targets = piece_attacks(p, f) & ~our & to_mask;
            while (targets) {
                int t = msb(targets);
                targets ^= bit(t);
                emit(make_move(f, t));
            }
        }
        if (from_mask & p.bb[KING]) castling_moves(from_mask, to_mask);
        u64 pawns = p.bb[PAWN] & our & from_mask;
        if (!pawns) return;
        u64 capturers = pawns;
        while (capturers) {
            int f = msb(capturers);
            capturers ^= bit(f);
            u64 targets = pawn_attacks(p.turn, f) & them & to_mask;
            while (targets) {
                int t = msb(targets);
                targets ^= bit(t);
                promos_or_plain(f, t);
            }
        }
        u64 single, dbl;
        if (p.turn == WHITE) {
            single = (pawns << 8) & ~occ;
            dbl = (single << 8) & ~occ & (RANK_1 << 24);
        } else {
            single = (pawns >> 8) & ~occ;
            dbl = (single >> 8) & ~occ & (RANK_1 << 32);
        }
        single &= to_mask;
        dbl &= to_mask;
        int back = p.turn == WHITE ? -8 : 8;
        while (single) {
            int t = msb(single);
            single ^= bit(t);
            promos_or_plain(t + back, t);
        }
        while (dbl) {
            int t = msb(dbl);
            dbl ^= bit(t);
            emit(make_move(t + 2 * back, t));
        }
        if (p.ep >= 0) ep_moves(from_mask, to_mask);
    }
    CB_HD void evasions(u64 checkers, u64 from_mask, u64 to_mask) {
        u64 sliders = checkers & (p.bb[BISHOP] | p.bb[ROOK] | p.bb[QUEEN]);
        u64 attacked = 0;
        while (sliders) {
            int c = msb(sliders);
            sliders ^= bit(c);
            attacked |= line_through(king, c) & ~bit(c);
        }
        if (bit(king) & from_mask) {
            u64 targets = king_attacks(king) & ~p.occ[p.turn] & ~attacked & to_mask;
            while (targets) {
                int t = msb(targets);
                targets ^= bit(t);
                emit(make_move(king, t));
            }
        }
        int checker = msb(checkers);
        if (bit(checker) == checkers) {
            u64 target = between(king, checker) | checkers;
            pseudo_legal(~p.bb[KING] & from_mask, target & to_mask);
            if (p.ep >= 0 && !(bit(p.ep) & target)) {
                int last_double = p.ep + (p.turn == WHITE ? -8 : 8);
                if (last_double == checker) ep_moves(from_mask, to_mask);
            }
        }
    }
};

CB_HDC void generate_legal(const Pos& p, MoveList& out, u64 from_mask = ~0ULL, u64 to_mask = ~0ULL) {
    out.n = 0;
    int king = king_square(p, p.turn);
    if (king < 0) {
        Gen g(p, out, -1, 0);
        g.pseudo_legal(from_mask, to_mask);
        return;
    }
    Gen g(p, out, king, slider_blockers(p, king));
    u64 checkers = attackers_mask(p, p.turn ^ 1, king, occupied(p));
    if (checkers) g.evasions(checkers, from_mask, to_mask);
    else g.pseudo_legal(from_mask, to_mask);
}
CB_HDC bool has_legal_moves(const Pos& p) {
    MoveList ml;
    generate_legal(p, ml);
    return ml.n > 0;
}
CB_HDC bool has_legal_en_passant(const Pos& p) {
    if (p.ep < 0) return false;
    // cheap reject: no own pawn attacks the ep square
    if (!(p.bb[PAWN] & p.occ[p.turn] & pawn_attacks(p.turn ^ 1, p.ep))) return false;
    MoveList ml;
    generate_legal(p, ml, ~0ULL, bit(p.ep));
    for (int i = 0; i < ml.n; ++i)
        if (is_en_passant(p, ml.m[i])) return true;
    return false;
}

// ---------------------------------------------------------------------------------- key / hash
CB_HD u64 mix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
CB_HD u64 key_hash(const Pos& p) {
    u64 h = 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < 6; ++i) h = mix64(h ^ p.bb[i]) + 0x9E3779B97F4A7C15ULL * (i + 1);
    h = mix64(h ^ p.occ[WHITE]);
    h = mix64(h ^ ((u64)p.turn | ((u64)p.castling << 8) | ((u64)(uint8_t)p.ep_key << 16)));
    return h;
}
// full transposition-key equality (python-chess Board._transposition_key)
CB_HD bool key_equal(const Pos& a, const Pos& b) {
    if (a.hash != b.hash) return false;
    for (int i = 0; i < 6; ++i)
        if (a.bb[i] != b.bb[i]) return false;
    return a.occ[0] == b.occ[0] && a.occ[1] == b.occ[1] && a.turn == b.turn && a.castling == b.castling &&
           a.ep_key == b.ep_key;
}

// ---------------------------------------------------------------------------------- make move
CB_HD void remove_piece(Pos& p, int sq) {
    u64 m = ~bit(sq);
    for (int i = 0; i < 6; ++i) p.bb[i] &= m;
    p.occ[0] &= m;
    p.occ[1] &= m;
}
CB_HD void set_piece(Pos& p, int sq, int type, int color) {
    remove_piece(p, sq);
    p.bb[type] |= bit(sq);
    p.occ[color] |= bit(sq);
}
// Board after `m` (python-chess Board.push).  Fills everything except `occur`, which needs the record
// before it (see count_occurrences).  `m` must be legal.
CB_HDC Pos push_move(const Pos& prev, Move m) {
    Pos p = prev;
    int f = move_from(m), t = move_to(m), promo = move_promo(m);
    int us = prev.turn;
    bool zeroing = is_zeroing(prev, m);
    p.halfmove = zeroing ? 0 : (uint16_t)(prev.halfmove + 1);
    p.ply = (uint16_t)(prev.ply + 1);
    int type = piece_type_at(prev, f);
    remove_piece(p, f);
    bool captured = (occupied(prev) & bit(t)) != 0;
    // castling rights: squares touched lose their right; a king move loses both
    uint8_t cr = prev.castling;
    u64 touched = bit(f) | bit(t);
    if (touched & bit(7)) cr &= ~CR_WK;
    if (touched & bit(0)) cr &= ~CR_WQ;
    if (touched & bit(63)) cr &= ~CR_BK;
    if (touched & bit(56)) cr &= ~CR_BQ;
    if (type == KING) cr &= us == WHITE ? ~(CR_WK | CR_WQ) : ~(CR_BK | CR_BQ);
    p.castling = cr;
    p.ep = -1;
    bool castle = false;
    if (type == KING) {
        int d = (t & 7) - (f & 7);
        castle = d > 1 || d < -1;
    }
    if (type == PAWN) {
        int d = t - f;
        if (d == 16 && (f >> 3) == 1) p.ep = (int8_t)(f + 8);
        else if (d == -16 && (f >> 3) == 6) p.ep = (int8_t)(f - 8);
        else if (t == prev.ep && (d == 7 || d == 9 || d == -7 || d == -9) && !captured)
            remove_piece(p, prev.ep + (us == WHITE ? -8 : 8));
    }
    if (promo) type = promo - 1;
    if (castle) {
        bool a_side = (t & 7) < (f & 7);
        int base = us == WHITE ? 0 : 56;
        remove_piece(p, base + (a_side ? 0 : 7));
        set_piece(p, t, KING, us);
        set_piece(p, base + (a_side ? 3 : 5), ROOK, us);
    } else {
        set_piece(p, t, type, us);
    }
    p.turn = (uint8_t)(us ^ 1);
    p.last_move = m;
    // is_irreversible(m) on the pre-move board: zeroing, castling rights reduced, or a legal ep existed
    p.irrev_in = (uint8_t)(zeroing || cr != prev.castling || prev.ep_key >= 0);
    p.ep_key = has_legal_en_passant(p) ? p.ep : (int8_t)-1;
    p.hash = key_hash(p);
    p.occur = 1;
    p.pad_[0] = p.pad_[1] = p.pad_[2] = 0;
    return p;
}

// ---------------------------------------------------------------------------------- game end
CB_HD bool has_insufficient_material(const Pos& p, int color) {
    u64 ours = p.occ[color];
    if (ours & (p.bb[PAWN] | p.bb[ROOK] | p.bb[QUEEN])) return false;
    if (ours & p.bb[KNIGHT])
        return popcnt(ours) <= 2 && !(p.occ[color ^ 1] & ~p.bb[KING] & ~p.bb[QUEEN]);
    if (ours & p.bb[BISHOP]) {
        bool same = !(p.bb[BISHOP] & DARK_SQUARES) || !(p.bb[BISHOP] & LIGHT_SQUARES);
        return same && !p.bb[PAWN] && !p.bb[KNIGHT];
    }
    return true;
}
CB_HD bool is_insufficient_material(const Pos& p) {
    return has_insufficient_material(p, WHITE) && has_insufficient_material(p, BLACK);
}

// The positions before `cur` in the same game are reached through `before` (the position one ply
// back, or nullptr) and `prev(q)` (the position before q, or nullptr past the start of what is
// stored).  Scans stop at the first position whose incoming move was irreversible, exactly like
// python-chess pops until is_irreversible(move).
template <class PrevFn>
CB_HDN int count_occurrences(const Pos& cur, const Pos* before, PrevFn prev) {
    int n = 1;
    const Pos* walk = &cur;
    const Pos* q = before;
    while (!walk->irrev_in && q) {
        if (key_equal(*q, cur) && n < 255) ++n;
        walk = q;
        q = prev(q);
    }
    return n;
}

// Threefold claim-ahead for ONE legal move of `cur`: does it reach a position that already occurred twice since
// the last irreversible move?  (python-chess can_claim_threefold_repetition, the "any legal move" part.)
template <class PrevFn>
CB_HDN bool move_claims_threefold(const Pos& cur, Move m, const Pos* before, PrevFn prev) {
    if (is_zeroing(cur, m)) return false;  // only reversible moves can repeat (others change material, pawns or castling rights)
    Pos nxt = push_move(cur, m);
    if (nxt.irrev_in) return false;
    int n = 0, k = 1;
    const Pos* walk = &cur;
    const Pos* q = before;
    while (!walk->irrev_in && q) {  // same side to move as nxt: 1, 3, 5, ... plies back
        if ((k & 1) && key_equal(*q, nxt)) {
            if (++n >= 2) return true;
        }
        walk = q;
        q = prev(q);
        ++k;
    }
    return false;
}

// outcome(claim_draw=True) (game.py:57-58) for a position whose `occur` is filled in, EXCEPT the threefold
// claim-ahead scan over the legal moves: *scan is set when that scan (move_claims_threefold for every legal
// move, any order) still has to decide between OUT_THREEFOLD_REPETITION and the returned OUT_NONE.
// `legal` / `n_legal` must hold generate_legal(cur).  max_len: config.max_game_length (game.py:55), 0 to disable.
// the part of outcome_before_scan that needs no history beyond cur.occur; OUT_NONE here means "not over unless the claim-ahead scan says so"
CB_HDN int outcome_before_history(const Pos& cur, const Move* legal, int n_legal, int max_len) {
    if (max_len > 0 && cur.ply == max_len) return OUT_MAX_LENGTH;
    if (n_legal == 0 && checkers_mask(cur)) return OUT_CHECKMATE;
    if (is_insufficient_material(cur)) return OUT_INSUFFICIENT_MATERIAL;
    if (n_legal == 0) return OUT_STALEMATE;
    if (cur.halfmove >= 150) return OUT_SEVENTYFIVE_MOVES;
    if (cur.occur >= 5) return OUT_FIVEFOLD_REPETITION;
    if (cur.halfmove >= 100) return OUT_FIFTY_MOVES;  // can_claim_fifty_moves
    if (cur.halfmove >= 99) {
        for (int i = 0; i < n_legal; ++i) {
            if (is_zeroing(cur, legal[i])) continue;
            Pos nxt = push_move(cur, legal[i]);
            if (has_legal_moves(nxt)) return OUT_FIFTY_MOVES;
        }
    }
    if (cur.occur >= 3) return OUT_THREEFOLD_REPETITION;  // can_claim_threefold_repetition
    return OUT_NONE;
}
template <class PrevFn>
CB_HDN int outcome_before_scan(const Pos& cur, const Move* legal, int n_legal, const Pos* before, PrevFn prev, int max_len, bool* scan) {
    *scan = false;
    if (max_len > 0 && cur.ply == max_len) return OUT_MAX_LENGTH;
    if (n_legal == 0 && checkers_mask(cur)) return OUT_CHECKMATE;
    if (is_insufficient_material(cur)) return OUT_INSUFFICIENT_MATERIAL;
    if (n_legal == 0) return OUT_STALEMATE;
    if (cur.halfmove >= 150) return OUT_SEVENTYFIVE_MOVES;
    if (cur.occur >= 5) return OUT_FIVEFOLD_REPETITION;
    // can_claim_fifty_moves
    if (cur.halfmove >= 100) return OUT_FIFTY_MOVES;
    if (cur.halfmove >= 99) {
        for (int i = 0; i < n_legal; ++i) {
            if (is_zeroing(cur, legal[i])) continue;
            Pos nxt = push_move(cur, legal[i]);
            if (has_legal_moves(nxt)) return OUT_FIFTY_MOVES;
        }
    }
    // can_claim_threefold_repetition
    if (cur.occur >= 3) return OUT_THREEFOLD_REPETITION;
    // claim-ahead needs a record that reaches at least three plies back without crossing an irreversible move
    if (cur.irrev_in || !before || before->irrev_in) return OUT_NONE;
    const Pos* b2 = prev(before);
    if (!b2 || b2->irrev_in || !prev(b2)) return OUT_NONE;
    *scan = true;
    return OUT_NONE;
}

template <class PrevFn>
CB_HDN int outcome_claim_draw(const Pos& cur, const MoveList& legal, const Pos* before, PrevFn prev, int max_len) {
    bool scan;
    const int out = outcome_before_scan(cur, legal.m, legal.n, before, prev, max_len, &scan);
    if (!scan) return out;
    for (int i = 0; i < legal.n; ++i)
        if (move_claims_threefold(cur, legal.m[i], before, prev)) return OUT_THREEFOLD_REPETITION;
    return OUT_NONE;
}
// terminal_value(to_play) of game.py:61-77 for the side to move at a terminal position
CB_HD int terminal_value_to_play(int outcome) { return outcome == OUT_CHECKMATE ? -1 : 0; }

// ---------------------------------------------------------------------------------- action index
// actionspace.uci_to_action (actionspace.py:116-180) in closed form: (file'*8 + rank')*73 + plane,
// coordinates from the mover's point of view.  Returns -1 where the reference returns None.
CB_HD int move_to_action(Move m, int turn) {
    int f = move_from(m), t = move_to(m), promo = move_promo(m);
    int fc = f & 7, fr = f >> 3, tc = t & 7, tr = t >> 3;
    if (turn == BLACK) { fc = 7 - fc; fr = 7 - fr; tc = 7 - tc; tr = 7 - tr; }
    int dx = tc - fc, dy = tr - fr;
    int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
    int plane = -1;
    if (adx * ady == 2) {
        // (1,2),(-1,2),(2,1),(2,-1),(1,-2),(-1,-2),(-2,1),(-2,-1) -> 56..63
        if (dy == 2) plane = dx == 1 ? 56 : 57;
        else if (dx == 2) plane = dy == 1 ? 58 : 59;
        else if (dy == -2) plane = dx == 1 ? 60 : 61;
        else plane = dy == 1 ? 62 : 63;
    } else if (promo && dy == 1 && adx <= 1) {
        int dir3 = dx == 0 ? 0 : (dx == 1 ? 1 : 2);  // N, NE, NW
        if (promo == 5) plane = dx == 0 ? 0 : (dx == 1 ? 1 : 7);
        else plane = 64 + dir3 * 3 + (promo == 4 ? 0 : (promo == 3 ? 1 : 2));  // R, B, N
    } else if ((dx == 0 || dy == 0 || adx == ady) && (dx || dy)) {
        int dist = adx > ady ? adx : ady;
        int sx = (dx > 0) - (dx < 0), sy = (dy > 0) - (dy < 0);
        // N,NE,E,SE,S,SW,W,NW
        int dir = sx == 0 ? (sy > 0 ? 0 : 4) : (sx > 0 ? (sy > 0 ? 1 : (sy == 0 ? 2 : 3)) : (sy < 0 ? 5 : (sy == 0 ? 6 : 7)));
        plane = 8 * (dist - 1) + dir;
    }
    if (plane < 0) return -1;
    return (fc * 8 + fr) * 73 + plane;
}
// order of UCI strings (mcts.py:120-123 breaks ucb ties by the larger move string):
// chars are file, rank, file, rank, promo with '' < 'b' < 'n' < 'q' < 'r'
CB_HD int uci_sort_key(Move m) {
    int f = move_from(m), t = move_to(m), promo = move_promo(m);
    int pk = promo == 0 ? 0 : (promo == 3 ? 1 : (promo == 2 ? 2 : (promo == 5 ? 3 : 4)));
    return ((((f & 7) * 8 + (f >> 3)) * 8 + (t & 7)) * 8 + (t >> 3)) * 8 + pk;
}

}  // namespace cb
