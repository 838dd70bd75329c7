// Warp-cooperative legal move generation (device only): the same moves in the same python-chess order as
// generate_legal (rules.cuh), but every from-square (or push target) is expanded by its own lane and the
// per-lane move runs are stitched together with a warp prefix sum.  One lane running the serial generator
// was the critical path of mcts_step_kernel (~36,000 of ~68,000 cycles per tree and step).
//
// Order reproduced: in check, king steps first, then — for a single checker — the
// blocks / captures; otherwise non-pawn pieces by from-square descending (targets descending), castling
// (h-side first), pawn captures (capturer descending, target descending, promotions q r b n), single pushes
// (target descending, promotions q r b n), double pushes (target descending), en passant (capturer descending).
#pragma once
#include "rules.cuh"

namespace cb {

__device__ __forceinline__ int nth_msb(u64 m, int i) {  // square of the i-th highest set bit, i < popcount(m)
    for (int k = 0; k < i; ++k) m ^= bit(msb(m));
    return msb(m);
}
__device__ __forceinline__ int warp_exclusive_scan(int v, int lane, int* total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}
// legal subset of `targets` for moves from f (no promotion piece: legality does not depend on it)
__device__ __forceinline__ u64 safe_targets(const Pos& p, int king, u64 blockers, int f, u64 targets) {
    u64 ok = 0;
    while (targets) {
        const int t = msb(targets);
        targets ^= bit(t);
        if (is_safe(p, king, blockers, make_move(f, t))) ok |= bit(t);
    }
    return ok;
}

// All 32 lanes call with the same position; `out` (shared memory, MAX_MOVES entries) receives the ordered legal
// moves, the count is returned to every lane.  The caller synchronises the warp before reading `out`.
__device__ int generate_legal_warp(const Pos& p, Move* out, int lane) {
    const int us = p.turn;
    const u64 our = p.occ[us], them = p.occ[us ^ 1], occ = our | them;
    const int king = king_square(p, us);
    if (king < 0) {  // no king: pseudo-legal moves, never reached from a real game; keep the serial generator
        int n = 0;
        if (lane == 0) {
            MoveList ml;
            generate_legal(p, ml);
            for (int i = 0; i < ml.n; ++i) out[i] = ml.m[i];
            n = ml.n;
        }
        return __shfl_sync(0xffffffffu, n, 0);
    }
    const u64 blockers = slider_blockers(p, king);
    const u64 checkers = attackers_mask(p, us ^ 1, king, occ);
    u64 from_mask = ~0ULL, to_mask = ~0ULL;
    int total = 0;
    bool ep_late = false;  // evasions: en passant that captures the checking pawn, generated after everything else
    if (checkers) {
        // ---- king steps (first in python-chess' evasion generator)
        u64 sliders = checkers & (p.bb[BISHOP] | p.bb[ROOK] | p.bb[QUEEN]), attacked = 0;
        while (sliders) {
            const int c = msb(sliders);
            sliders ^= bit(c);
            attacked |= line_through(king, c) & ~bit(c);
        }
        const u64 ktargets = king_attacks(king) & ~our & ~attacked;
        const int nk = popcnt(ktargets);
        int t = -1;
        bool ok = false;
        if (lane < nk) {
            t = nth_msb(ktargets, lane);
            ok = is_safe(p, king, blockers, make_move(king, t));
        }
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (ok) out[__popc(okm & ((1u << lane) - 1))] = make_move(king, t);
        total = __popc(okm);
        const int checker = msb(checkers);
        if (bit(checker) != checkers) return total;  // double check: only the king moves
        const u64 target = between(king, checker) | checkers;
        from_mask = ~p.bb[KING];
        to_mask = target;
        if (p.ep >= 0 && !(bit(p.ep) & target) && p.ep + (us == WHITE ? -8 : 8) == checker) ep_late = true;
    }
    // ---- pass A: non-pawn pieces on lanes 0..15, castling on lane 16
    {
        const u64 pieces = our & ~p.bb[PAWN] & from_mask;
        const int np = popcnt(pieces);
        int f = -1, cnt = 0;
        u64 legal = 0;
        Move castle[2];
        if (lane < np) {
            f = nth_msb(pieces, lane);
            legal = safe_targets(p, king, blockers, f, piece_attacks(p, f) & ~our & to_mask);
            cnt = popcnt(legal);
        } else if (lane == 16 && !checkers) {
            MoveList ml;  // castling is rare and at most two moves: the serial generator, restricted to the king
            ml.n = 0;
            Gen g(p, ml, king, blockers);
            g.castling_moves(~0ULL, ~0ULL);
            cnt = ml.n;
            castle[0] = ml.n > 0 ? ml.m[0] : 0;
            castle[1] = ml.n > 1 ? ml.m[1] : 0;
        }
        int sum;
        int o = total + warp_exclusive_scan(cnt, lane, &sum);
        if (lane < np) {
            while (legal) {
                const int t = msb(legal);
                legal ^= bit(t);
                out[o++] = make_move(f, t);
            }
        } else if (lane == 16) {
            for (int i = 0; i < cnt; ++i) out[o++] = castle[i];
        }
        total += sum;
    }
    // ---- pass B: pawn captures on lanes 0..7, single pushes 8..15, double pushes 16..23, en passant 24
    {
        const u64 pawns = p.bb[PAWN] & our & from_mask;
        u64 single, dbl;
        if (us == WHITE) {
            single = (pawns << 8) & ~occ;
            dbl = (single << 8) & ~occ & (RANK_1 << 24);
        } else {
            single = (pawns >> 8) & ~occ;
            dbl = (single >> 8) & ~occ & (RANK_1 << 32);
        }
        single &= to_mask;
        dbl &= to_mask;
        const int back = us == WHITE ? -8 : 8;
        const int grp = lane >> 3, idx = lane & 7;
        int f = -1, cnt = 0;
        u64 legal = 0;
        Move epm[2];
        if (grp == 0 && idx < popcnt(pawns)) {
            f = nth_msb(pawns, idx);
            legal = safe_targets(p, king, blockers, f, pawn_attacks(us, f) & them & to_mask);
            cnt = popcnt(legal & ~(RANK_1 | RANK_8)) + 4 * popcnt(legal & (RANK_1 | RANK_8));
        } else if (grp == 1 && idx < popcnt(single)) {
            const int t = nth_msb(single, idx);
            f = t + back;
            if (is_safe(p, king, blockers, make_move(f, t))) { legal = bit(t); cnt = (bit(t) & (RANK_1 | RANK_8)) ? 4 : 1; }
        } else if (grp == 2 && idx < popcnt(dbl)) {
            const int t = nth_msb(dbl, idx);
            f = t + 2 * back;
            if (is_safe(p, king, blockers, make_move(f, t))) { legal = bit(t); cnt = 1; }
        } else if (lane == 24 && p.ep >= 0) {
            MoveList ml;
            ml.n = 0;
            Gen g(p, ml, king, blockers);
            if (ep_late) g.ep_moves(~0ULL, ~0ULL);
            else g.ep_moves(from_mask, to_mask);
            cnt = ml.n;
            epm[0] = ml.n > 0 ? ml.m[0] : 0;
            epm[1] = ml.n > 1 ? ml.m[1] : 0;
        }
        int sum;
        int o = total + warp_exclusive_scan(cnt, lane, &sum);
        if (lane == 24) {
            for (int i = 0; i < cnt; ++i) out[o++] = epm[i];
        } else {
            while (legal) {
                const int t = msb(legal);
                legal ^= bit(t);
                if (bit(t) & (RANK_1 | RANK_8)) {  // promotions q, r, b, n (double pushes never land here)
                    out[o++] = make_move(f, t, 5); out[o++] = make_move(f, t, 4); out[o++] = make_move(f, t, 3); out[o++] = make_move(f, t, 2);
                } else {
                    out[o++] = make_move(f, t);
                }
            }
        }
        total += sum;
    }
    return total;
}

}  // namespace cb
