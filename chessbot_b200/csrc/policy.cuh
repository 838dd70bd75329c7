// K5 — legal-move gather + exp + normalise (mcts.expand, mcts.py:91-101) for one position per warp.
//
// priors:  p_i = exp(logit[action_i]) (f32), s = np.sum(p, dtype=f32), P_i = p_i / s  — bit-exact with
// NumPy: the exponential is a 65,536-entry table indexed by the bf16 logit and filled by the host
// (np.exp on that host, so identical by construction; NumPy's f32 exp is not correctly rounded and
// differs between CPUs), and the sum reproduces NumPy's pairwise reduction order.
#pragma once
#include "common.cuh"
#include "rules.cuh"

namespace cb {

// np.sum(float32[n], dtype=float32) for n <= 256 (NumPy pairwise_sum: n < 8 sequential; blocks of
// <= 128 with eight interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and a
// sequential tail; n > 128 splits at n/2 rounded down to a multiple of 8).  `v` is in shared memory.
// All lanes must call; every lane returns the sum.
__device__ __forceinline__ float numpy_sum_f32_warp(const float* v, int n, int lane) {
    float res = 0.f;
    if (n < 8) {
        if (lane == 0) {
            res = n > 0 ? v[0] : 0.f;
            for (int i = 1; i < n; ++i) res = __fadd_rn(res, v[i]);
        }
        return __shfl_sync(0xffffffffu, res, 0);
    }
    int h = 0;  // split point: [0,h) and [h,n); h = 0 means a single block
    if (n > 128) {
        h = n / 2;
        h -= h % 8;
    }
    // lanes 0..7 own the accumulators of block 0, lanes 8..15 those of block 1
    const int blk = lane >> 3, j = lane & 7;
    float r = 0.f;
    int lo = 0, len = 0;
    if (blk == 0) { lo = 0; len = h ? h : n; }
    else if (blk == 1 && h) { lo = h; len = n - h; }
    if (len >= 8) {
        r = v[lo + j];
        const int lim = len - (len % 8);
        for (int i = 8; i < lim; i += 8) r = __fadd_rn(r, v[lo + i + j]);
    }
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) within each group of 8 lanes
    float t = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1));   // valid at j = 0,2,4,6
    t = __fadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2));         // valid at j = 0,4
    t = __fadd_rn(t, __shfl_down_sync(0xffffffffu, t, 4));         // valid at j = 0
    if (j == 0 && len >= 8) {
        for (int i = len - (len % 8); i < len; ++i) t = __fadd_rn(t, v[lo + i]);
    }
    const float s0 = __shfl_sync(0xffffffffu, t, 0), s1 = __shfl_sync(0xffffffffu, t, 8);
    return h ? __fadd_rn(s0, s1) : s0;
}

}  // namespace cb
