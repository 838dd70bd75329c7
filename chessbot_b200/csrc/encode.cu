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
#include "encode.cuh"

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

    // ---- phase 1 + 2a (encode.cuh: the same warp routines mcts_step_kernel uses for the leaf it selects)
    encode_masks_warp(pos_base, fi, s_mask[warp], s_scalar[warp], lane);
    if (planes_act) encode_cells_warp(s_mask[warp], s_scalar[warp], pos_i, planes_act, lane);
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
