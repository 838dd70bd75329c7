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

// ---------------------------------------------------------------------------------- repetition history, warp-cooperative
// python-chess's repetition scans (is_repetition, can_claim_threefold_repetition) walk the move stack back, one position per step, until
// a move was irreversible.  On the device the positions before a tree leaf are a linked list (tree path, then the root's game record):
// walked serially that is one dependent memory access per ply, three times per new leaf (occurrence count + the claim-ahead scan in
// two rounds of 32 legal moves), up to 100 plies deep - the tail that set mcts_step's duration.  Here the whole window is fetched in
// parallel, 32 plies per round: the tree part through the recorded path (node -> position id), the record part by index (a record is
// contiguous, one ply per slot), and key hash + position id land in shared memory; the scans then compare hashes there and touch a
// full position only on a hash match.
constexpr int MCTS_HIST = 128;

// W = number of positions before `cur` (1 .. W plies back) those scans examine: q_a is examined iff cur and q_1 .. q_(a-1) were all
// entered by reversible moves and q_a is stored.  h_hash[a-1] / h_pos[a-1] are filled for a <= W.  -1: window longer than MCTS_HIST.
// path[0 .. depth-1]: nodes root .. leaf; the leaf's own position is `cur`.
__device__ int gather_history_warp(const Pos* pool, const Node* nodes, const int32_t* path, int depth, int root_pos, int rec_start, bool cur_irrev,
                                   u64* h_hash, int32_t* h_pos, int lane) {
    if (cur_irrev) return 0;
    int W = 0;
    for (int base = 0; base < MCTS_HIST; base += 32) {
        const int a = base + lane + 1;  // plies back
        int pid = -1;
        if (a <= depth - 1) {
            pid = nodes[path[depth - 1 - a]].pos_id;
        } else {
            const int idx = root_pos - (a - (depth - 1));
            if (idx >= rec_start) pid = idx;
        }
        u64 h = 0;
        bool irrev = false;
        if (pid >= 0) {
            h = pool[pid].hash;
            irrev = pool[pid].irrev_in != 0;
        }
        h_hash[a - 1] = h;
        h_pos[a - 1] = pid;
        const unsigned exist = __ballot_sync(0xffffffffu, pid >= 0);
        const unsigned irr = __ballot_sync(0xffffffffu, pid >= 0 && irrev);
        const int n_exist = exist == 0xffffffffu ? 32 : __ffs(~exist) - 1;  // the stored positions are a prefix of the round
        const int first_irr = irr ? __ffs(irr) - 1 : 32;                     // the position entered by an irreversible move is still examined
        const int take = min(n_exist, first_irr + 1);
        W += take;
        if (take < 32) return W;
    }
    return -1;
}

// count_occurrences(cur, ...) over the gathered window: every lane returns the count (capped at 255 like the serial version)
__device__ int count_occurrences_warp(const Pos& cur, const Pos* pool, const u64* h_hash, const int32_t* h_pos, int W, int lane) {
    int n = 1;
    for (int base = 0; base < W; base += 32) {
        const int i = base + lane;
        const bool hit = i < W && h_hash[i] == cur.hash && key_equal(pool[h_pos[i]], cur);
        n += __popc(__ballot_sync(0xffffffffu, hit));
    }
    return n < 255 ? n : 255;
}

// move_claims_threefold(cur, m, ...) over the gathered window (one lane per move)
__device__ bool move_claims_threefold_hist(const Pos& cur, Move m, const Pos* pool, const u64* h_hash, const int32_t* h_pos, int W) {
    if (is_zeroing(cur, m)) return false;
    Pos nxt = push_move(cur, m);
    if (nxt.irrev_in) return false;
    int n = 0;
    for (int a = 1; a <= W; a += 2)  // same side to move as nxt: 1, 3, 5, ... plies back
        if (h_hash[a - 1] == nxt.hash && key_equal(pool[h_pos[a - 1]], nxt)) {
            if (++n >= 2) return true;
        }
    return false;
}

}  // namespace cb
