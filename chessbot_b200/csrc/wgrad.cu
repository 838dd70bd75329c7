// T3 — weight gradient of a 3x3 "same" convolution (train.py:150-164 -> model.fit; Conv2D kernels of model.py:19-37)
// as split-K implicit GEMMs on tcgen05 tensor cores:
//
//   dW[tap][ci][co] = sum over boards b and squares (r, c) of  A[b, r + dy, c + dx, ci] * dY[b, r, c, co]
//
// * Both operands are the tile-native activation buffers as they lie in HBM (common.cuh).  The reduction runs over
//   squares, and a 16-byte cell holds 8 CHANNELS of one square, so both operands are "MN-major" for the tensor core:
//   the no-swizzle descriptor takes SBO = 3200 B (next 8-channel chunk) and LBO = 160 B (next group of 8 squares);
//   one MMA consumes K = 16 squares = one board row of both boards of the pair.  The nine taps are nine start
//   addresses into the input tile (its zero border is the "same" padding), exactly as in the forward kernel.
// * A CTA owns one (block of 128 input channels, half of the output channels, row of three taps): three accumulators of
//   128 x N/2 f32 in TMEM (<= 384 columns), summed over its share of the board pairs (split-K over the batch).
// * The two CTAs of a cluster take the two output-channel halves of the same job and board pairs: the input tile is
//   fetched once per cluster (each CTA loads one 64-channel block and multicasts it to both), the dY halves are private.
//   A slot is refilled only after the MMAs of BOTH CTAs that read it have retired (tcgen05.commit multicast).
// * Partial sums go to a scratch buffer [split][tap][ci][co]; wgrad_reduce_kernel adds the splits in a fixed order
//   (bit-reproducible gradients), drops the channel padding and folds the stem's remainder channels (encode.cu).
#include "kernels.cuh"

namespace cb {

constexpr int WG_THREADS = 192;  // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (one per TMEM lane quarter)
constexpr int WG_STAGES = 2;
constexpr int WG_A_BYTES = 2 * ACT_K64_BYTES;  // 128 input channels of a board pair
constexpr int WG_B_BYTES = 2 * ACT_K64_BYTES;  // up to 128 output channels of the pair
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
struct WgradSmem {
    uint64_t full[WG_STAGES], empty[WG_STAGES], acc_full;
    uint32_t tmem_base;
};
constexpr size_t WG_SMEM_BYTES = 1024 + (size_t)WG_STAGES * WG_STAGE_BYTES;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG_THREADS, 1) wgrad_tcgen05_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    WgradSmem* s = reinterpret_cast<WgradSmem*>(smem);
    uint8_t* stages = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cid = blockIdx.x >> 1;
    const int job = cid / p.splits, split = cid % p.splits;  // job = input-channel block * 3 + tap row
    const int ci_block = job / 3, trow = job % 3;
    const int nh = p.n >> 1;  // output channels of this CTA
    const int lo = (int)((long long)p.pairs * split / p.splits), hi = (int)((long long)p.pairs * (split + 1) / p.splits);
    const int a_blocks = p.cj_in >= 16 ? 2 : 1;  // 64-channel blocks in the input tile (a 64-channel layer has one)
    const int dy_chunks = nh >> 3;               // 8-channel chunks of this CTA's dY half: 16, 8 or 4
    const int dy_copies = (dy_chunks + 7) >> 3;
    const uint32_t stage_tx = (uint32_t)(a_blocks * ACT_K64_BYTES + dy_chunks * ACT_CHUNK_BYTES);

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { ptx::mbar_init(&s->full[i], 1); ptx::mbar_init(&s->empty[i], 2); }
        ptx::mbar_init(&s->acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&s->tmem_base, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // the peer multicasts into this CTA's shared memory and arrives on its barriers
    ptx::tc_fence_after();
    const uint32_t tmem = s->tmem_base;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) { ptx::tma_prefetch_desc(&p.tm_a); ptx::tma_prefetch_desc(&p.tm_dy); }
        for (int it = lo; it < hi; ++it) {
            const int k = it - lo, st = k % WG_STAGES;
            ptx::mbar_wait(&s->empty[st], ((k / WG_STAGES) & 1) ^ 1);
            if (ptx::elect_one()) {
                uint8_t* a_dst = stages + st * WG_STAGE_BYTES;
                uint8_t* b_dst = a_dst + WG_A_BYTES;
                ptx::mbar_expect_tx(&s->full[st], stage_tx);
                // rows of 128 B: one 8-channel chunk of a board pair = 25 rows
                if (a_blocks == 2)
                    ptx::tma_load_2d_multicast(a_dst + rank * ACT_K64_BYTES, &p.tm_a, 0, (it * p.cj_in + 16 * ci_block + 8 * (int)rank) * 25,
                                               &s->full[st], 3);
                else if (rank == 0)
                    ptx::tma_load_2d_multicast(a_dst, &p.tm_a, 0, it * p.cj_in * 25, &s->full[st], 3);
                for (int c = 0; c < dy_copies; ++c)
                    ptx::tma_load_2d(b_dst + c * ACT_K64_BYTES, &p.tm_dy, 0, (it * (p.n >> 3) + (int)rank * dy_chunks + 8 * c) * 25, &s->full[st]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        const uint32_t idesc = ptx::idesc_bf16_f32_mn(128, nh);
        for (int it = lo; it < hi; ++it) {
            const int k = it - lo, st = k % WG_STAGES;
            ptx::mbar_wait(&s->full[st], (k / WG_STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t a_base = ptx::smem_u32(stages + st * WG_STAGE_BYTES), b_base = a_base + WG_A_BYTES;
            if (ptx::elect_one()) {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {  // K = 16 squares: board row ks of both boards
                        const uint64_t a_desc = ptx::smem_desc(a_base + (uint32_t)(trow * 20 + t + ks * 20) * 16, 160, ACT_CHUNK_BYTES);
                        const uint64_t b_desc = ptx::smem_desc(b_base + (uint32_t)(21 + ks * 20) * 16, 160, ACT_CHUNK_BYTES);
                        ptx::umma_bf16(tmem + (uint32_t)(t * nh), a_desc, b_desc, idesc, (k | ks) != 0);
                    }
                }
                ptx::umma_commit_mc(&s->empty[st], 3);  // the slot is free in BOTH CTAs when these MMAs retire (its input half is shared)
            }
            __syncwarp();
        }
        if (hi > lo && ptx::elect_one()) ptx::umma_commit(&s->acc_full);
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue: TMEM -> partial sums in HBM
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int ci = ci_block * 128 + m;
        const int ci_pad = p.cj_in * 8;
        const bool valid = ci < ci_pad;
        if (hi > lo) {
            ptx::mbar_wait(&s->acc_full, 0);
            ptx::tc_fence_after();
        }
        for (int t = 0; t < 3; ++t) {
            const int tap = trow * 3 + t;
            float* dst = p.partial + (((size_t)split * 9 + tap) * ci_pad + ci) * p.n + rank * nh;
            for (int cb = 0; cb < (nh >> 5); ++cb) {
                uint32_t acc[32];
                if (hi > lo) {
                    ptx::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * nh + cb * 32), acc);
                    ptx::tmem_wait_ld();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[e] = 0u;
                }
                if (valid) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<uint4*>(dst + cb * 32 + e) = make_uint4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // the peer may still be multicasting into this CTA / signalling its barriers
    if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// dW[tap][ci][co] (HWIO, real channel counts) = sum of the split partials in a fixed order.  Stem: input channels 119 / 120
// of the planes carry the bf16 remainders of channels 113 / 118 (encode.cu), their products belong to those channels.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int ci_pad, int n, int cin, int c, int stem,
                                    float* __restrict__ out) {
    const size_t total = (size_t)9 * cin * c;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx % c), ci = (int)((idx / c) % cin), tap = (int)(idx / ((size_t)c * cin));
        const int extra = stem ? (ci == 113 ? 119 : (ci == 118 ? 120 : -1)) : -1;
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) {
            const float* base = partial + ((size_t)s * 9 + tap) * ci_pad * n;
            acc += base[(size_t)ci * n + co];
            if (extra >= 0) acc += base[(size_t)extra * n + co];
        }
        out[idx] = acc;
    }
}

// Plain CUDA-core restatement on the same bf16 inputs (fp32 accumulate, padded layout [9][ci_pad][n]).  Test infrastructure
// for the tcgen05 kernel (cb_debug_wgrad with use_reference); never on the product path.
__global__ void wgrad_reference_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ dy, int pairs, int cj_in, int n,
                                       float* __restrict__ out) {
    const int ci_pad = cj_in * 8, cjo = n >> 3;
    const size_t total = (size_t)9 * ci_pad * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx % n), ci = (int)((idx / n) % ci_pad), tap = (int)(idx / ((size_t)n * ci_pad));
        const int dy_ = tap / 3 - 1, dx_ = tap % 3 - 1;
        float acc = 0.f;
        for (int pair = 0; pair < pairs; ++pair)
            for (int bip = 0; bip < 2; ++bip)
                for (int sq = 0; sq < 64; ++sq) {
                    const int row = sq >> 3, col = sq & 7;
                    const float a = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(
                        reinterpret_cast<const uint8_t*>(in) + act_offset_bytes(pair, cj_in, ci >> 3, act_cell(row + dy_, bip, col + dx_)) + (ci & 7) * 2));
                    const float g = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(
                        reinterpret_cast<const uint8_t*>(dy) + act_offset_bytes(pair, cjo, co >> 3, act_cell(row, bip, col)) + (co & 7) * 2));
                    acc = fmaf(a, g, acc);
                }
        out[idx] = acc;
    }
}

int wgrad_splits(int num_sms, int cj_in) {
    const int jobs = (cj_in >= 32 ? 2 : 1) * 3;  // clusters per split
    const int s = (num_sms / 2) / jobs;
    return s < 1 ? 1 : s;
}

int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
    static std::atomic<uint64_t> smem_set;
    if (cb_opt_in_dynamic_smem((const void*)wgrad_tcgen05_kernel, (int)WG_SMEM_BYTES, smem_set)) return -1;
    if ((p.n != 64 && p.n != 128 && p.n != 256) || (p.cj_in != 8 && p.cj_in != 16 && p.cj_in != 32) || p.pairs < 1 || p.splits < 1) {
        cb_set_error("wgrad: channels (padded) must be 64, 128 or 256");
        return -1;
    }
    const int jobs = (p.cj_in >= 32 ? 2 : 1) * 3;
    wgrad_tcgen05_kernel<<<2 * jobs * p.splits, WG_THREADS, WG_SMEM_BYTES, stream>>>(p);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_wgrad_reduce(const float* partial, int splits, int ci_pad, int n, int cin, int c, int stem, float* out, cudaStream_t stream) {
    const size_t total = (size_t)9 * cin * c;
    wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(partial, splits, ci_pad, n, cin, c, stem, out);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_wgrad_reference(const __nv_bfloat16* in, const __nv_bfloat16* dy, int pairs, int cj_in, int n, float* out, cudaStream_t stream) {
    wgrad_reference_kernel<<<592, 256, 0, stream>>>(in, dy, pairs, cj_in, n, out);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cb
