// K1 — board -> input planes (gameimage.board_to_image, gameimage.py:149-192) as a warp-per-position
// bitboard expansion.
//
// Input: up to 8 frames (newest first) of 96-byte game-record positions (rules.cuh: Pos), addressed
// through an index list so the same kernel serves the stand-alone API and the tree (frames found by
// walking parents).  Output: bf16 planes in the tile-native layout of the convolution (common.cuh),
// 128 channels: 0..111 = 8 frames x 14 planes, 112..118 = colour, move count, castling x4, halfmove
// clock, 119/120 = the parts of the move count / halfmove clock that do not fit bf16's 8-bit mantissa
// (the stem convolution carries duplicated weights for them, so integer inputs stay exact), 121..127 = 0.
// Optionally the reference's own int16 [8,8,119] image for bit-exact parity checks.
//
// Phase 1: the warp turns the frames into 112 oriented 64-bit masks in shared memory (orientation is
// per frame: White to move -> row = 7 - rank, col = file; Black to move -> row = rank, col = 7 - file,
// "player 1" = the side to move of that frame, gameimage.py:37-41).  Phase 2: each lane writes 16-byte
// cells (8 channels of one square), 512 contiguous-ish bytes per warp instruction.
#include "kernels.cuh"

namespace cb {

constexpr int ENC_WARPS = 4;

__global__ void __launch_bounds__(ENC_WARPS * 32) encode_planes_kernel(const Pos* __restrict__ pos_base,
                                                                        const int32_t* __restrict__ frame_idx,  // [n][8], newest first, -1 = absent
                                                                        int n, __nv_bfloat16* __restrict__ planes_act,
                                                                        int16_t* __restrict__ planes_i16,
                                                                        const int32_t* __restrict__ slot_list,  // compacted batch: position i
                                                                        const int* __restrict__ n_dev) {        // = frames of slot_list[i], i < *n_dev
    __shared__ u64 s_mask[ENC_WARPS][112];
    __shared__ int s_scalar[ENC_WARPS][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pos_i = blockIdx.x * ENC_WARPS + warp;
    if (pos_i >= n || (n_dev && pos_i >= *n_dev)) return;
    const int32_t* fi = frame_idx + (size_t)(slot_list ? slot_list[pos_i] : pos_i) * 8;

    // ---- phase 1: 112 oriented masks (channel c = 14*slot + plane, slot = 7 - t for t plies back)
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
        s_mask[warp][c] = m;
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
        s_scalar[warp][lane] = v;
    }
    __syncwarp();

    // ---- phase 2a: bf16 cells in tile-native layout.  Per 8-channel chunk the warp writes the 64
    // squares in two instructions; 8 lanes cover one board row = 128 contiguous bytes.
    if (planes_act) {
        const int pair = pos_i >> 1, bip = pos_i & 1;
        uint8_t* out = reinterpret_cast<uint8_t*>(planes_act);
        const int ply = s_scalar[warp][1], hm = s_scalar[warp][6];
        // what bf16 holds exactly (8 significant bits) stays in place, the remainder goes next door
        const float ply_hi = __uint_as_float(__float_as_uint((float)ply) & 0xFFFF0000u);
        const float hm_hi = __uint_as_float(__float_as_uint((float)hm) & 0xFFFF0000u);
        for (int chunk = 0; chunk < 16; ++chunk) {
            uint32_t w_scalar[4] = {0, 0, 0, 0};
            if (chunk >= 14) {
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (chunk == 14) {
#pragma unroll
                    for (int e = 0; e < 6; ++e) v[e] = (float)s_scalar[warp][e];
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
                        if ((s_mask[warp][chunk * 8 + e] >> sq) & 1) w[e >> 1] |= (e & 1) ? 0x3F800000u : 0x00003F80u;
                }
                uint8_t* dst = out + act_offset_bytes(pair, 16, chunk, act_cell(sq >> 3, bip, sq & 7));
                *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                // whole 32-B sectors per board row: the first / last column also store the zero border cell next to theirs
                if ((sq & 7) == 0) *reinterpret_cast<uint4*>(dst - 16) = make_uint4(0u, 0u, 0u, 0u);
                if ((sq & 7) == 7) *reinterpret_cast<uint4*>(dst + 16) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
    // ---- phase 2b: the reference's int16 [8][8][119] image (exact integers)
    if (planes_i16) {
        int16_t* o = planes_i16 + (size_t)pos_i * 64 * 119;
        for (int i = lane; i < 64 * 119; i += 32) {
            const int sq = i / 119, c = i % 119;
            o[i] = (int16_t)(c < 112 ? (int)((s_mask[warp][c] >> sq) & 1) : s_scalar[warp][c - 112]);
        }
    }
}

int launch_encode(const Pos* pos_base, const int32_t* frame_idx, int n, __nv_bfloat16* planes_act, int16_t* planes_i16,
                  cudaStream_t stream, const int32_t* slot_list, const int* n_dev) {
    if (n <= 0) return 0;
    encode_planes_kernel<<<(n + ENC_WARPS - 1) / ENC_WARPS, ENC_WARPS * 32, 0, stream>>>(pos_base, frame_idx, n, planes_act, planes_i16, slot_list,
                                                                                    n_dev);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cb
