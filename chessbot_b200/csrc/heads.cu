// K3/K4 — value and policy heads (model.py:43-56) after the tower.
//
// The 1x1 head convolutions (C->1 and C->2) are fused into the last tower epilogue (conv.cu) and
// arrive here as raw per-square sums head_raw[board][64][3].  This file finishes them:
//   value : BN -> LeakyReLU -> Flatten(64) -> Dense(64->C) -> LeakyReLU -> Dense(C->1) -> tanh
//   policy: BN -> LeakyReLU -> Flatten(h,w,c = 128) -> Dense(128->4672)
// The policy Dense is evaluated either densely (predict_fn contract, model.py:73-80) or only at the
// legal moves of the position (search): both walk k = 0..127 with one fmaf chain per action, so a
// logit has the same bits whichever way it was produced.
#include "kernels.cuh"

namespace cb {


__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

// one block of 256 threads per board
__global__ void __launch_bounds__(256) head_features_kernel(HeadParams p, int boards) {
    __shared__ float v0[64];
    __shared__ float red[8];
    const int b = blockIdx.x;
    if (b >= boards || (p.n_boards_dev && b >= *p.n_boards_dev)) return;
    const int dst = p.slot_list ? p.slot_list[b] : b;  // compacted batch: results go back to the leaf's evaluation slot
    const float* raw = p.head_raw + (size_t)b * 192;
    const int t = threadIdx.x;
    if (t < 64) v0[t] = lrelu(fmaf(raw[t * 3], p.bn[0], p.bn[1]), p.slope);
    if (t < 128) {
        const int sq = t >> 1, c = t & 1;
        p.pfeat[(size_t)dst * 128 + t] = lrelu(fmaf(raw[sq * 3 + 1 + c], p.bn[2 + 2 * c], p.bn[3 + 2 * c]), p.slope);
    }
    __syncthreads();
    float part = 0.f;
    if (t < p.c) {
        float h = 0.f;
        for (int sq = 0; sq < 64; ++sq) h = fmaf(v0[sq], p.fc1[sq * p.c + t], h);
        part = lrelu(h, p.slope) * p.fc2[t];
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((t & 31) == 0) red[t >> 5] = part;
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        p.value[dst] = tanhf(s);
    }
}

// dense policy logits: grid (ceil(4672/256), ceil(boards/8)), thread = one action for 8 boards
__global__ void __launch_bounds__(256) policy_dense_kernel(const float* __restrict__ pfeat, const float* __restrict__ w /*[128][4672]*/,
                                                            int boards, __nv_bfloat16* __restrict__ logits_bf16,
                                                            float* __restrict__ logits_f32) {
    __shared__ float pf[8][128];
    const int b0 = blockIdx.y * 8;
    for (int i = threadIdx.x; i < 8 * 128; i += 256) {
        const int bb = i >> 7;
        pf[bb][i & 127] = (b0 + bb < boards) ? pfeat[(size_t)(b0 + bb) * 128 + (i & 127)] : 0.f;
    }
    __syncthreads();
    const int a = blockIdx.x * 256 + threadIdx.x;
    if (a >= 4672) return;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 128; ++k) {
        const float wk = w[(size_t)k * 4672 + a];
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) acc[bb] = fmaf(pf[bb][k], wk, acc[bb]);
    }
    for (int bb = 0; bb < 8 && b0 + bb < boards; ++bb) {
        const size_t o = (size_t)(b0 + bb) * 4672 + a;
        if (logits_bf16) logits_bf16[o] = __float2bfloat16_rn(acc[bb]);
        if (logits_f32) logits_f32[o] = acc[bb];
    }
}

int launch_head_features(const HeadParams& p, int boards, cudaStream_t stream) {
    if (p.c > 256) { cb_set_error("heads: filters > 256"); return -1; }
    head_features_kernel<<<boards, 256, 0, stream>>>(p, boards);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_policy_dense(const float* pfeat, const float* w, int boards, __nv_bfloat16* logits_bf16, float* logits_f32,
                        cudaStream_t stream) {
    dim3 grid((4672 + 255) / 256, (boards + 7) / 8);
    policy_dense_kernel<<<grid, 256, 0, stream>>>(pfeat, w, boards, logits_bf16, logits_f32);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------ synthetic network
// Device twin of oracle/synthetic_net.py: the mock-network contract of the reference's tests
// (tests/mcts_v2.py:51-67) as an integer hash of the input planes.  Reads the exact int16 image.
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
__global__ void __launch_bounds__(128) synth_hash_kernel(const int16_t* __restrict__ planes_i16, int n, uint32_t seed,
                                                         uint32_t* __restrict__ hash_out, float* __restrict__ value) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + warp;
    if (i >= n) return;
    const int16_t* img = planes_i16 + (size_t)i * 7616;
    uint32_t h = 0;
    for (int k = lane; k < 7616; k += 32) h += (uint32_t)(int)img[k] * ((uint32_t)k * 2654435761u + 0x9E3779B9u);
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    h ^= seed;
    if (lane == 0) {
        hash_out[i] = h;
        value[i] = (float)((int)((fmix32(h) >> 8) & 0xFFFF) - 32768) / 32768.0f;
    }
}
__global__ void synth_logits_dense_kernel(const uint32_t* __restrict__ hash, int n, __nv_bfloat16* __restrict__ logits) {
    const size_t total = (size_t)n * 4672;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t a = (uint32_t)(idx % 4672);
        const uint32_t u = fmix32(hash[idx / 4672] ^ ((a + 1u) * 0x9E3779B1u));
        logits[idx] = __float2bfloat16_rn((float)((int)((u >> 16) & 0xFF) - 128) / 32.0f);
    }
}
int launch_synth(const int16_t* planes_i16, int n, uint32_t seed, uint32_t* hash_out, float* value, __nv_bfloat16* logits_dense,
                 cudaStream_t stream) {
    synth_hash_kernel<<<(n + 3) / 4, 128, 0, stream>>>(planes_i16, n, seed, hash_out, value);
    CB_CUDA_OK(cudaGetLastError());
    if (logits_dense) {
        synth_logits_dense_kernel<<<296, 256, 0, stream>>>(hash_out, n, logits_dense);
        CB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace cb
