// K1 as a warp routine (device only): board -> bf16 input planes of ONE position in the convolution's tile-native layout, by the 32 lanes of
// the calling warp.  Shared by encode_planes_kernel (stand-alone API, root positions) and mcts_step_kernel, which encodes the leaf it has
// just selected instead of leaving it to a separate launch.  See encode.cu for the plane layout and orientation rules (gameimage.py:149-192).
#pragma once
#include "kernels.cuh"

namespace cb {

// Phase 1: the frames (newest first, -1 = absent) become 112 oriented 64-bit masks + 8 scalars in the warp's shared-memory scratch.
__device__ __forceinline__ void encode_masks_warp(const Pos* __restrict__ pos_base, const int32_t* __restrict__ fi, u64* s_mask /*[112]*/,
                                                  int* s_scalar /*[8]*/, int lane) {
    for (int c = lane; c < 112; c += 32) {
        const int slot = c / 14, plane = c % 14, t = 7 - slot;
        const int idx = fi[t];
        u64 m = 0;
        if (idx >= 0) {
            const Pos& f = pos_base[idx];
            const int turn = f.turn;
            if (plane < 12) {
                // plane order P,N,B,R,K,Q (gameimage.py:10-17); bb order P,N,B,R,Q,K
                const int pt = plane % 6;
                const int bbi = pt == 4 ? KING : (pt == 5 ? QUEEN : pt);
                const int color = plane < 6 ? turn : (turn ^ 1);
                const u64 raw = f.bb[bbi] & f.occ[color];
                // White to move: index = (7-rank)*8 + file = sq ^ 56 (byte swap).
                // Black to move: index = rank*8 + (7-file) = sq ^ 7 (bit-reverse then byte swap).
                const u64 x = turn == WHITE ? raw : __brevll(raw);
                m = ((u64)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
            } else {
                m = (f.occur >= (plane == 12 ? 2 : 3)) ? ~0ULL : 0ULL;  // is_repetition(2 | 3), gameimage.py:72-74
            }
        }
        s_mask[c] = m;
    }
    if (lane < 8) {
        const Pos& cur = pos_base[fi[0]];
        const int w = cur.turn == WHITE;
        int v = 0;
        switch (lane) {
            case 0: v = w; break;                                                      // colour (gameimage.py:88-91)
            case 1: v = cur.ply; break;                                                // len(move_stack) (:103)
            case 2: v = (cur.castling & (w ? CR_WK : CR_BK)) != 0; break;              // mover kingside (:119-132)
            case 3: v = (cur.castling & (w ? CR_WQ : CR_BQ)) != 0; break;
            case 4: v = (cur.castling & (w ? CR_BK : CR_WK)) != 0; break;              // opponent
            case 5: v = (cur.castling & (w ? CR_BQ : CR_WQ)) != 0; break;
            case 6: v = cur.halfmove; break;                                           // halfmove clock (:146)
            default: v = 0;
        }
        s_scalar[lane] = v;
    }
    __syncwarp();
}

// Phase 2a: bf16 cells of board `board` in the tile-native layout.  Per 8-channel chunk the warp writes the 64 squares in two
// instructions; 8 lanes cover one board row = 128 contiguous bytes.
__device__ __forceinline__ void encode_cells_warp(const u64* s_mask, const int* s_scalar, int board, __nv_bfloat16* __restrict__ planes_act, int lane) {
    const int pair = board >> 1, bip = board & 1;
    uint8_t* out = reinterpret_cast<uint8_t*>(planes_act);
    const int ply = s_scalar[1], hm = s_scalar[6];
    // what bf16 holds exactly (8 significant bits) stays in place, the remainder goes next door
    const float ply_hi = __uint_as_float(__float_as_uint((float)ply) & 0xFFFF0000u);
    const float hm_hi = __uint_as_float(__float_as_uint((float)hm) & 0xFFFF0000u);
    for (int chunk = 0; chunk < 16; ++chunk) {
        uint32_t w_scalar[4] = {0, 0, 0, 0};
        if (chunk >= 14) {
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (chunk == 14) {
#pragma unroll
                for (int e = 0; e < 6; ++e) v[e] = (float)s_scalar[e];
                v[1] = ply_hi;
                v[6] = hm_hi;
                v[7] = (float)ply - ply_hi;  // channel 119
            } else {
                v[0] = (float)hm - hm_hi;    // channel 120
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v[e], v[e + 1]);
                w_scalar[e >> 1] = *reinterpret_cast<uint32_t*>(&b2);
            }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int sq = half * 32 + lane;  // row*8 + col
            uint32_t w[4] = {w_scalar[0], w_scalar[1], w_scalar[2], w_scalar[3]};
            if (chunk < 14) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if ((s_mask[chunk * 8 + e] >> sq) & 1) w[e >> 1] |= (e & 1) ? 0x3F800000u : 0x00003F80u;
            }
            uint8_t* dst = out + act_offset_bytes(pair, 16, chunk, act_cell(sq >> 3, bip, sq & 7));
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            // whole 32-B sectors per board row: the first / last column also store the zero border cell next to theirs
            if ((sq & 7) == 0) *reinterpret_cast<uint4*>(dst - 16) = make_uint4(0u, 0u, 0u, 0u);
            if ((sq & 7) == 7) *reinterpret_cast<uint4*>(dst + 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

}  // namespace cb
