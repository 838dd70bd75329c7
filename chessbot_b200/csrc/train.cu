// T1/T2/T4..T9 — the element-wise, head and optimizer kernels of the training step (train.py:150-164 -> model.fit on the graph
// of model.py:19-70): BatchNormalization in training mode (batch statistics, moving averages), LeakyReLU and the residual add,
// both heads, the two losses, their gradients, and SGD with Nesterov momentum + L2.
//
// The three dense contractions of a layer run on tensor cores elsewhere: forward and data gradient through the tower kernel
// (conv.cu, one layer per launch with an identity epilogue: the raw convolution is needed for the batch statistics), the
// weight gradient through wgrad.cu.  Everything here is HBM-bound streaming over the tile-native activation buffers
// (common.cuh): one thread per 16-byte cell = 8 channels of one square, a block walks one 8-channel chunk over a range of
// board pairs, so a thread keeps its 8 channels' statistics / scale / shift in registers.
#include "kernels.cuh"

namespace cb {

namespace {
__device__ __forceinline__ void unpack8(const uint4& v, float (&x)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[2 * i] = __uint_as_float(w[i] << 16);
        x[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&x)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 t = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&t);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ size_t cell_off(int pair, int cj, int j, int cell) { return ((size_t)(pair * cj + j) * ACT_CELLS + cell) * 16; }

// one thread per SQUARE of a board pair (128 threads): thread t -> row t / 16, board (t / 8) % 2, column t % 8; 8 consecutive
// threads touch 128 contiguous bytes.  The border cells (zero, never written) are skipped.
constexpr int SQ_THREADS = 128;
__device__ __forceinline__ int square_cell(int t) { return ((t >> 4) + 1) * 20 + ((t >> 3) & 1) * 10 + (t & 7) + 1; }

// A board row is 8 cells = 128 bytes that start 16 bytes into a 32-byte DRAM sector (the border cell comes first): storing only
// the squares leaves the first and last sector of every row partially written, and L2 has to fetch them from DRAM before it can
// write them back.  The threads of the first / last column therefore also store the (zero) border cell next to theirs: every row
// becomes five whole sectors.
__device__ __forceinline__ void store_row_cell(uint8_t* base, size_t off, int t, const uint4& v) {
    *reinterpret_cast<uint4*>(base + off) = v;
    if ((t & 7) == 0) *reinterpret_cast<uint4*>(base + off - 16) = make_uint4(0u, 0u, 0u, 0u);
    if ((t & 7) == 7) *reinterpret_cast<uint4*>(base + off + 16) = make_uint4(0u, 0u, 0u, 0u);
}

// sum over the 32 lanes of a warp of 32 values per lane, transposed: afterwards lane c holds the total of value c in v[0]
// (butterfly with halving payload: 31 shuffles instead of 32 x 5; the same helper lives in conv.cu for the BN statistics)
__device__ __forceinline__ void warp_transpose_sum32_t(float (&v)[32], int lane) {
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
        const bool up = (lane & k) != 0;
#pragma unroll
        for (int i = 0; i < k; ++i) {
            const float send = up ? v[i] : v[i + k];
            const float recv = __shfl_xor_sync(0xffffffffu, send, k);
            v[i] = (up ? v[i + k] : v[i]) + recv;
        }
    }
}

// sum `v[0..NV)` over the block, result in thread `i` for value i (i < NV) as double
template <int NV>
__device__ __forceinline__ double block_sum(float (&v)[NV], float* red /*[warps][NV]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp * NV + i] = x;
    }
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x < NV)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[w * NV + threadIdx.x];
    return s;
}
}  // namespace

// ------------------------------------------------------------------------------------------------ input
// sample images (train.py:104-110: int16 [B][8][8][119]) -> the stem's bf16 planes, 128 channels, tile-native: channels 113 /
// 118 (move count, halfmove clock) keep what bf16 holds exactly, the remainder goes to channels 119 / 120 (as encode.cu)
__global__ void planes_from_images_kernel(const int16_t* __restrict__ img, int n, uint8_t* __restrict__ planes) {
    const size_t total = (size_t)n * 64 * 16;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(idx & 15), sq = (int)((idx >> 4) & 63), b = (int)(idx >> 10);
        const int16_t* px = img + ((size_t)b * 64 + sq) * 119;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ch = chunk * 8 + e;
            v[e] = ch < 119 ? (float)px[ch] : 0.f;
        }
        if (chunk == 14) {
            const float ply = (float)px[113], hm = (float)px[118];
            const float ply_hi = __uint_as_float(__float_as_uint(ply) & 0xFFFF0000u), hm_hi = __uint_as_float(__float_as_uint(hm) & 0xFFFF0000u);
            v[1] = ply_hi;
            v[6] = hm_hi;
            v[7] = ply - ply_hi;
        } else if (chunk == 15) {
            const float hm = (float)px[118];
            v[0] = hm - __uint_as_float(__float_as_uint(hm) & 0xFFFF0000u);
        }
        store_row_cell(planes, cell_off(b >> 1, 16, chunk, act_cell(sq >> 3, b & 1, sq & 7)), sq, pack8(v));
    }
}

// ------------------------------------------------------------------------------------------------ BatchNormalization, forward
// grid (cj, splits): per-channel sum and sum of squares of the raw convolution over the batch (border cells are zero)
__global__ void __launch_bounds__(SQ_THREADS) bn_stats_kernel(const uint8_t* __restrict__ y, int pairs, int cj, double* __restrict__ stats) {
    __shared__ float red[4 * 16];
    const int j = blockIdx.x, t = threadIdx.x, cell = square_cell(t);
    const int lo = (int)((long long)pairs * blockIdx.y / gridDim.y), hi = (int)((long long)pairs * (blockIdx.y + 1) / gridDim.y);
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    {
#pragma unroll 4
        for (int p = lo; p < hi; ++p) {
            float x[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(y + cell_off(p, cj, j, cell))), x);
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc[e] += x[e]; acc[8 + e] = fmaf(x[e], x[e], acc[8 + e]); }
        }
    }
    const double s = block_sum<16>(acc, red);
    if (t < 16) atomicAdd(stats + (t >> 3) * cj * 8 + j * 8 + (t & 7), s);
}

// out = LeakyReLU(gamma * (y - mean) * rstd + beta (+ skip)); saves mean / rstd for the backward pass and updates the moving
// statistics (Keras: moving = moving * momentum + batch * (1 - momentum), biased batch variance)
__global__ void __launch_bounds__(SQ_THREADS) bn_apply_kernel(const uint8_t* __restrict__ y, const uint8_t* __restrict__ skip, uint8_t* __restrict__ out,
                                                        int n, int pairs, int cj, int c, const double* __restrict__ stats,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ mean_rstd, float* __restrict__ moving, float eps, float slope,
                                                        float momentum) {
    __shared__ float s_scale[8], s_shift[8];
    const int j = blockIdx.x, t = threadIdx.x, cpad = cj * 8;
    const int lo = (int)((long long)pairs * blockIdx.y / gridDim.y), hi = (int)((long long)pairs * (blockIdx.y + 1) / gridDim.y);
    if (t < 8) {  // the chunk's 8 channels: statistics -> scale / shift once per block (double: sum of squares minus squared sum)
        const int ch = j * 8 + t;
        float sc = 0.f, sh = 0.f;
        if (ch < c) {
            const double count = (double)n * 64.0;
            const double mean = stats[ch] / count;
            double var = stats[cpad + ch] / count - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const double rstd = 1.0 / sqrt(var + (double)eps);
            sc = (float)((double)gamma[ch] * rstd);
            sh = (float)((double)beta[ch] - mean * (double)gamma[ch] * rstd);
            if (blockIdx.y == 0) {
                mean_rstd[ch] = (float)mean;
                mean_rstd[cpad + ch] = (float)rstd;
                mean_rstd[2 * cpad + ch] = sc;  // the backward pass recomputes the sign of a skip-less layer's pre-activation from y
                mean_rstd[3 * cpad + ch] = sh;
                moving[ch] = moving[ch] * momentum + (float)mean * (1.f - momentum);
                moving[c + ch] = moving[c + ch] * momentum + (float)var * (1.f - momentum);
            }
        }
        s_scale[t] = sc;
        s_shift[t] = sh;
    }
    __syncthreads();
    float scale[8], shift[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { scale[e] = s_scale[e]; shift[e] = s_shift[e]; }
    const int cell = square_cell(t), bip = (t >> 3) & 1;
#pragma unroll 4
    for (int p = lo; p < hi; ++p) {
        if (2 * p + bip >= n) continue;  // phantom board of an odd batch stays zero
        const size_t off = cell_off(p, cj, j, cell);
        float x[8], s[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(y + off)), x);
        if (skip) unpack8(__ldg(reinterpret_cast<const uint4*>(skip + off)), s);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float v = fmaf(x[e], scale[e], shift[e]);
            if (skip) v += s[e];
            x[e] = v > 0.f ? v : v * slope;
        }
        store_row_cell(out, off, t, pack8(x));
    }
}

// ------------------------------------------------------------------------------------------------ BatchNormalization, backward
// dz = g * LeakyReLU'(a);  sums[ch] = sum dz, sums[cpad + ch] = sum dz * xhat   (xhat = (y - mean) * rstd)
__global__ void __launch_bounds__(SQ_THREADS, 8) bn_bwd_reduce_kernel(const uint8_t* __restrict__ g, const uint8_t* __restrict__ a, const uint8_t* __restrict__ y,
                                                             int n, int pairs, int cj, const float* __restrict__ mean_rstd, float slope,
                                                             double* __restrict__ sums) {
    __shared__ float red[4 * 16];
    const int j = blockIdx.x, t = threadIdx.x, cpad = cj * 8;
    const int lo = (int)((long long)pairs * blockIdx.y / gridDim.y), hi = (int)((long long)pairs * (blockIdx.y + 1) / gridDim.y);
    float mean[8], rstd[8], sc[8], sh[8], acc[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        mean[e] = mean_rstd[j * 8 + e]; rstd[e] = mean_rstd[cpad + j * 8 + e];
        sc[e] = mean_rstd[2 * cpad + j * 8 + e]; sh[e] = mean_rstd[3 * cpad + j * 8 + e];
        acc[e] = acc[8 + e] = 0.f;
    }
    {
        const int cell = square_cell(t), bip = (t >> 3) & 1;
#pragma unroll 2
        for (int p = lo; p < hi; ++p) {
            if (2 * p + bip >= n) continue;
            const size_t off = cell_off(p, cj, j, cell);
            float gv[8], av[8], yv[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(g + off)), gv);
            unpack8(__ldg(reinterpret_cast<const uint4*>(y + off)), yv);
            if (a) unpack8(__ldg(reinterpret_cast<const uint4*>(a + off)), av);
            else {  // no residual add: the activation's sign is the sign of scale * y + shift, as the forward pass computed it
#pragma unroll
                for (int e = 0; e < 8; ++e) av[e] = fmaf(yv[e], sc[e], sh[e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float dz = av[e] > 0.f ? gv[e] : gv[e] * slope;
                acc[e] += dz;
                acc[8 + e] = fmaf(dz, (yv[e] - mean[e]) * rstd[e], acc[8 + e]);
            }
        }
    }
    const double s = block_sum<16>(acc, red);
    if (t < 16) atomicAdd(sums + (t >> 3) * cpad + j * 8 + (t & 7), s);
}

// dy = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat)); optionally also stores dz (the residual branch's gradient);
// dgamma = sum dz * xhat, dbeta = sum dz
__global__ void __launch_bounds__(SQ_THREADS, 8) bn_bwd_apply_kernel(const uint8_t* __restrict__ g, const uint8_t* __restrict__ a, const uint8_t* __restrict__ y,
                                                            uint8_t* __restrict__ dy, uint8_t* __restrict__ dz_out, int n, int pairs, int cj, int c,
                                                            const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                            const double* __restrict__ sums, float slope, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
    const int j = blockIdx.x, t = threadIdx.x, cpad = cj * 8;
    const int lo = (int)((long long)pairs * blockIdx.y / gridDim.y), hi = (int)((long long)pairs * (blockIdx.y + 1) / gridDim.y);
    const double inv_count = 1.0 / ((double)n * 64.0);
    float mean[8], rstd[8], sc[8], sh[8], k0[8], m1[8], m2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int ch = j * 8 + e;
        mean[e] = mean_rstd[ch];
        rstd[e] = mean_rstd[cpad + ch];
        sc[e] = mean_rstd[2 * cpad + ch];
        sh[e] = mean_rstd[3 * cpad + ch];
        const bool real = ch < c;
        k0[e] = real ? gamma[ch] * rstd[e] : 0.f;
        m1[e] = (float)(sums[ch] * inv_count);
        m2[e] = (float)(sums[cpad + ch] * inv_count);
        if (real && blockIdx.y == 0 && t == e) { dgamma[ch] = (float)sums[cpad + ch]; dbeta[ch] = (float)sums[ch]; }
    }
    const int cell = square_cell(t), bip = (t >> 3) & 1;
#pragma unroll 2
    for (int p = lo; p < hi; ++p) {
        if (2 * p + bip >= n) continue;
        const size_t off = cell_off(p, cj, j, cell);
        float gv[8], av[8], yv[8], o[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(g + off)), gv);
        unpack8(__ldg(reinterpret_cast<const uint4*>(y + off)), yv);
        if (a) unpack8(__ldg(reinterpret_cast<const uint4*>(a + off)), av);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) av[e] = fmaf(yv[e], sc[e], sh[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float dz = av[e] > 0.f ? gv[e] : gv[e] * slope;
            gv[e] = dz;
            o[e] = k0[e] * (dz - m1[e] - (yv[e] - mean[e]) * rstd[e] * m2[e]);
        }
        store_row_cell(dy, off, t, pack8(o));
        if (dz_out) store_row_cell(dz_out, off, t, pack8(gv));
    }
}

// ------------------------------------------------------------------------------------------------ heads, forward
// the three 1x1 head convolutions (model.py:43,52): head_raw[board][square][3] = (value, policy 0, policy 1); block = one board pair
__global__ void __launch_bounds__(128) head_conv_fwd_kernel(const uint8_t* __restrict__ a, int n, int cj, const float* __restrict__ wv /*[c]*/,
                                                             const float* __restrict__ wp /*[c][2]*/, int c, float* __restrict__ head_raw) {
    __shared__ float w[3][256];
    const int t = threadIdx.x, pair = blockIdx.x;
    for (int i = t; i < cj * 8; i += 128) {
        w[0][i] = i < c ? wv[i] : 0.f;
        w[1][i] = i < c ? wp[i * 2] : 0.f;
        w[2][i] = i < c ? wp[i * 2 + 1] : 0.f;
    }
    __syncthreads();
    const int bip = t >> 6, sq = t & 63, board = 2 * pair + bip;
    if (board >= n) return;
    const int cell = act_cell(sq >> 3, bip, sq & 7);
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
    for (int j = 0; j < cj; ++j) {
        float x[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(a + cell_off(pair, cj, j, cell))), x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            h0 = fmaf(x[e], w[0][j * 8 + e], h0);
            h1 = fmaf(x[e], w[1][j * 8 + e], h1);
            h2 = fmaf(x[e], w[2][j * 8 + e], h2);
        }
    }
    float* o = head_raw + ((size_t)board * 64 + sq) * 3;
    o[0] = h0; o[1] = h1; o[2] = h2;
}

// batch statistics of the three head channels: stats[h] = sum, stats[3 + h] = sum of squares
__global__ void __launch_bounds__(256) head_stats_kernel(const float* __restrict__ head_raw, int n, double* __restrict__ stats) {
    __shared__ float red[8 * 6];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const size_t total = (size_t)n * 64;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const float x = head_raw[i * 3 + h];
            acc[h] += x;
            acc[3 + h] = fmaf(x, x, acc[3 + h]);
        }
    }
    const double s = block_sum<6>(acc, red);
    if (threadIdx.x < 6) atomicAdd(stats + threadIdx.x, s);
}

// per board: head BN (batch statistics) + LeakyReLU -> value features [64] and policy features [128] (Flatten order h, w, c);
// Dense(64 -> C) -> LeakyReLU -> Dense(C -> 1) -> tanh.  Keeps what the backward pass needs.  hbn = value.bn gamma, beta (2) then
// policy.bn gamma[2], beta[2]; hmoving = value mean, var, policy mean[2], var[2].
__global__ void __launch_bounds__(256) head_features_train_kernel(const float* __restrict__ head_raw, int n, const double* __restrict__ stats,
                                                                   const float* __restrict__ vbn, const float* __restrict__ pbn,
                                                                   float* __restrict__ vmoving, float* __restrict__ pmoving,
                                                                   const float* __restrict__ fc1, const float* __restrict__ fc2, int c, float eps,
                                                                   float slope, float momentum, float* __restrict__ hmr /*[6] mean x3, rstd x3*/,
                                                                   float* __restrict__ vfeat, float* __restrict__ pfeat, float* __restrict__ h1pre,
                                                                   float* __restrict__ value) {
    __shared__ float v0[64];
    __shared__ float red[8];
    __shared__ float sc[3], sh[3];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < 3) {
        const double count = (double)n * 64.0;
        const double mean = stats[t] / count;
        double var = stats[3 + t] / count - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const float gamma = t == 0 ? vbn[0] : pbn[t - 1], beta = t == 0 ? vbn[1] : pbn[2 + t - 1];
        sc[t] = (float)(gamma * rstd);
        sh[t] = (float)(beta - mean * gamma * rstd);
        if (b == 0) {
            hmr[t] = (float)mean;
            hmr[3 + t] = (float)rstd;
            float* mm = t == 0 ? vmoving : pmoving + (t - 1);
            float* mv = t == 0 ? vmoving + 1 : pmoving + 2 + (t - 1);
            *mm = *mm * momentum + (float)mean * (1.f - momentum);
            *mv = *mv * momentum + (float)var * (1.f - momentum);
        }
    }
    __syncthreads();
    const float* raw = head_raw + (size_t)b * 192;
    if (t < 64) {
        const float x = fmaf(raw[t * 3], sc[0], sh[0]);
        v0[t] = x > 0.f ? x : x * slope;
        vfeat[(size_t)b * 64 + t] = v0[t];
    }
    if (t < 128) {
        const int sq = t >> 1, k = t & 1;
        const float x = fmaf(raw[sq * 3 + 1 + k], sc[1 + k], sh[1 + k]);
        pfeat[(size_t)b * 128 + t] = x > 0.f ? x : x * slope;
    }
    __syncthreads();
    float part = 0.f;
    if (t < c) {
        float h = 0.f;
        for (int sq = 0; sq < 64; ++sq) h = fmaf(v0[sq], fc1[sq * c + t], h);
        h1pre[(size_t)b * c + t] = h;
        part = (h > 0.f ? h : h * slope) * fc2[t];
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((t & 31) == 0) red[t >> 5] = part;
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        value[b] = tanhf(s);
    }
}

// C[m][n] = sum_k opA(A)[m][k] * opB(B)[k][n], row-major f32; ta: A is stored [k][m]; tb: B is stored [n][k].  The dense layers
// of the heads (Dense 128 -> 4672, 64 -> C) and their gradients: plain GEMMs, 0.1 % of the step's FLOPs, on the legacy
// mma.sync path with TF32 inputs (10-bit mantissa, f32 accumulate): 128 x 128 x 16 tiles, 8 warps of 64 x 32.
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
template <int BM, int BN>  // 128 x 128 for the big products, 64 x 64 where the output is small and the reduction long (more CTAs)
__global__ void __launch_bounds__(256) gemm_tf32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                                                         int ta, int tb, int k_per_split) {
    // split-K: block z reduces k in [z * k_per_split, (z + 1) * k_per_split) into its own M x N slab of C (summed afterwards in a
    // fixed order); gridDim.z == 1: the whole product straight into C
    constexpr int BK = 32, LD = BK + 4;  // row stride 36 words: the 8 x 4 fragment reads hit 32 different banks
    constexpr int MI = BM / 32, NJ = BN / 32;  // 16 x 8 MMA tiles per warp (warps 2 x 4)
    __shared__ uint32_t As[BM][LD], Bs[BN][LD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 2) * (BM / 2), wn = (warp & 3) * (BN / 4);
    const int g = lane >> 2, t4 = lane & 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
    C += (size_t)blockIdx.z * M * N;
    float acc[MI][NJ][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        // global -> shared in float4 along the operand's contiguous dimension (all of them are multiples of 4 here)
        for (int i = threadIdx.x; i < BM * BK / 4; i += 256) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ta) {  // A stored [k][m]
                const int mm = (i % (BM / 4)) * 4, kk = i / (BM / 4);
                const int gm = m0 + mm, gk = k0 + kk;
                if (gm < M && gk < k_end) v = *reinterpret_cast<const float4*>(A + (size_t)gk * M + gm);
                As[mm][kk] = to_tf32(v.x); As[mm + 1][kk] = to_tf32(v.y); As[mm + 2][kk] = to_tf32(v.z); As[mm + 3][kk] = to_tf32(v.w);
            } else {   // A stored [m][k]
                const int kk = (i % (BK / 4)) * 4, mm = i / (BK / 4);
                const int gm = m0 + mm, gk = k0 + kk;
                if (gm < M && gk < k_end) v = *reinterpret_cast<const float4*>(A + (size_t)gm * K + gk);
                *reinterpret_cast<uint4*>(&As[mm][kk]) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            }
        }
        for (int i = threadIdx.x; i < BN * BK / 4; i += 256) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tb) {  // B stored [n][k]
                const int kb = (i % (BK / 4)) * 4, nn = i / (BK / 4);
                const int gn = n0 + nn, gk = k0 + kb;
                if (gn < N && gk < k_end) v = *reinterpret_cast<const float4*>(B + (size_t)gn * K + gk);
                *reinterpret_cast<uint4*>(&Bs[nn][kb]) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            } else {   // B stored [k][n]
                const int nn = (i % (BN / 4)) * 4, kb = i / (BN / 4);
                const int gn = n0 + nn, gk = k0 + kb;
                if (gn < N && gk < k_end) v = *reinterpret_cast<const float4*>(B + (size_t)gk * N + gn);
                Bs[nn][kb] = to_tf32(v.x); Bs[nn + 1][kb] = to_tf32(v.y); Bs[nn + 2][kb] = to_tf32(v.z); Bs[nn + 3][kb] = to_tf32(v.w);
            }
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < BK; ks += 8) {
            uint32_t af[MI][4], bf[NJ][2];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int r = wm + i * 16 + g;
                af[i][0] = As[r][ks + t4]; af[i][1] = As[r + 8][ks + t4]; af[i][2] = As[r][ks + t4 + 4]; af[i][3] = As[r + 8][ks + t4 + 4];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int cc = wn + j * 8 + g;
                bf[j][0] = Bs[cc][ks + t4]; bf[j][1] = Bs[cc][ks + t4 + 4];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[i][j][0]), "+f"(acc[i][j][1]), "+f"(acc[i][j][2]), "+f"(acc[i][j][3])
                                 : "r"(af[i][0]), "r"(af[i][1]), "r"(af[i][2]), "r"(af[i][3]), "r"(bf[j][0]), "r"(bf[j][1]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int r = m0 + wm + i * 16 + g, cc = n0 + wn + j * 8 + 2 * t4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gm = r + (e >> 1) * 8, gn = cc + (e & 1);
                if (gm < M && gn < N) C[(size_t)gm * N + gn] = acc[i][j][e];
            }
        }
}

// ------------------------------------------------------------------------------------------------ losses
// CategoricalCrossentropy(from_logits=True) against the visit distribution (model.py:67; train.py:13-19), mean over the batch;
// logits are replaced by d loss / d logits = (softmax * sum(pi) - pi) / n.  losses[1] += sum over boards.
__global__ void __launch_bounds__(256) softmax_ce_kernel(float* __restrict__ logits, const float* __restrict__ pi, int n, double* __restrict__ losses) {
    __shared__ float red[8];
    const int b = blockIdx.x, t = threadIdx.x;
    float* l = logits + (size_t)b * 4672;
    const float* p = pi + (size_t)b * 4672;
    auto reduce = [&](float v, bool is_max) -> float {
        for (int o = 16; o > 0; o >>= 1) {
            const float u = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_max ? fmaxf(v, u) : v + u;
        }
        __syncthreads();
        if ((t & 31) == 0) red[t >> 5] = v;
        __syncthreads();
        float r = red[0];
        for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
        return r;
    };
    float mx = -INFINITY;
    for (int a = t; a < 4672; a += 256) mx = fmaxf(mx, l[a]);
    mx = reduce(mx, true);
    float se = 0.f, sp = 0.f, dot = 0.f;
    for (int a = t; a < 4672; a += 256) {
        se += expf(l[a] - mx);
        sp += p[a];
        dot = fmaf(p[a], l[a], dot);
    }
    se = reduce(se, false);
    sp = reduce(sp, false);
    dot = reduce(dot, false);
    const float lse = mx + logf(se);
    if (t == 0) atomicAdd(losses + 1, (double)(sp * lse - dot));
    const float inv_n = 1.f / (float)n;
    for (int a = t; a < 4672; a += 256) l[a] = (expf(l[a] - lse) * sp - p[a]) * inv_n;
}

// mean_squared_error on the tanh value (model.py:66) and its gradient down to the first value Dense:
// dh1pre[b][c] = dL/dpre * fc2[c] * LeakyReLU'(h1pre);  dpre[b] kept for dfc2.  losses[0] += (v - z)^2.
__global__ void __launch_bounds__(256) value_loss_bwd_kernel(const float* __restrict__ value, const float* __restrict__ z, const float* __restrict__ h1pre,
                                                              const float* __restrict__ fc2, int n, int c, float slope, float* __restrict__ dpre,
                                                              float* __restrict__ dh1pre, double* __restrict__ losses) {
    const int b = blockIdx.x, t = threadIdx.x;
    const float v = value[b], d = v - z[b];
    const float dp = 2.f * d / (float)n * (1.f - v * v);
    if (t == 0) { atomicAdd(losses, (double)(d * d)); dpre[b] = dp; }
    if (t < c) dh1pre[(size_t)b * c + t] = dp * fc2[t] * (h1pre[(size_t)b * c + t] > 0.f ? 1.f : slope);
}
// dfc2[c] = sum_b LeakyReLU(h1pre[b][c]) * dpre[b]; grid (ceil(c / 64), 64 batch slices), out zeroed before
__global__ void __launch_bounds__(64) value_fc2_grad_kernel(const float* __restrict__ h1pre, const float* __restrict__ dpre, int n, int c, float slope,
                                                             float* __restrict__ dfc2) {
    const int ch = blockIdx.x * 64 + threadIdx.x;
    if (ch >= c) return;
    const int lo = (int)((long long)n * blockIdx.y / gridDim.y), hi = (int)((long long)n * (blockIdx.y + 1) / gridDim.y);
    float acc = 0.f;
    for (int b = lo; b < hi; ++b) {
        const float h = h1pre[(size_t)b * c + ch];
        acc = fmaf(h > 0.f ? h : h * slope, dpre[b], acc);
    }
    atomicAdd(dfc2 + ch, acc);
}

// ------------------------------------------------------------------------------------------------ heads, backward
// head BN backward, reduction: hsums[h] = sum dz_h, hsums[3 + h] = sum dz_h * xhat_h over all boards and squares
__global__ void __launch_bounds__(256) head_bn_bwd_reduce_kernel(const float* __restrict__ head_raw, const float* __restrict__ vfeat,
                                                                  const float* __restrict__ pfeat, const float* __restrict__ dvfeat,
                                                                  const float* __restrict__ dpfeat, int n, const float* __restrict__ hmr, float slope,
                                                                  double* __restrict__ hsums) {
    __shared__ float red[8 * 6];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const size_t total = (size_t)n * 64;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const float f = h == 0 ? vfeat[i] : pfeat[i * 2 + h - 1];
            const float df = h == 0 ? dvfeat[i] : dpfeat[i * 2 + h - 1];
            const float dz = f > 0.f ? df : df * slope;
            acc[h] += dz;
            acc[3 + h] = fmaf(dz, (head_raw[i * 3 + h] - hmr[h]) * hmr[3 + h], acc[3 + h]);
        }
    }
    const double s = block_sum<6>(acc, red);
    if (threadIdx.x < 6) atomicAdd(hsums + threadIdx.x, s);
}

// head BN backward (apply) + 1x1 convolutions backward: the tower gradient G[b, sq, ch] = sum_h draw_h * w_h[ch] (bf16, tile-native)
// and the head kernel gradients dw_h[ch] = sum draw_h * a[ch].  Block = `ppb` board pairs; dwv / dwp zeroed before.
__global__ void __launch_bounds__(128) head_conv_bwd_kernel(const uint8_t* __restrict__ a, uint8_t* __restrict__ gout, int n, int pairs, int cj, int c,
                                                             int ppb, const float* __restrict__ head_raw, const float* __restrict__ vfeat,
                                                             const float* __restrict__ pfeat, const float* __restrict__ dvfeat,
                                                             const float* __restrict__ dpfeat, const float* __restrict__ hmr,
                                                             const double* __restrict__ hsums, const float* __restrict__ vbn,
                                                             const float* __restrict__ pbn, const float* __restrict__ wv, const float* __restrict__ wp,
                                                             float slope, float* __restrict__ dwv, float* __restrict__ dwp,
                                                             float* __restrict__ dvbn, float* __restrict__ dpbn) {
    __shared__ float w[3][256];
    __shared__ float dw[3][256];
    const int t = threadIdx.x, lane = t & 31;
    for (int i = t; i < cj * 8; i += 128) {
        w[0][i] = i < c ? wv[i] : 0.f;
        w[1][i] = i < c ? wp[i * 2] : 0.f;
        w[2][i] = i < c ? wp[i * 2 + 1] : 0.f;
        dw[0][i] = dw[1][i] = dw[2][i] = 0.f;
    }
    __syncthreads();
    const double inv_count = 1.0 / ((double)n * 64.0);
    float k0[3], m1[3], m2[3];
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        const float gamma = h == 0 ? vbn[0] : pbn[h - 1];
        k0[h] = gamma * hmr[3 + h];
        m1[h] = (float)(hsums[h] * inv_count);
        m2[h] = (float)(hsums[3 + h] * inv_count);
    }
    if (blockIdx.x == 0 && t < 3) {  // dgamma = sum dz * xhat, dbeta = sum dz
        if (t == 0) { dvbn[0] = (float)hsums[3]; dvbn[1] = (float)hsums[0]; }
        else { dpbn[t - 1] = (float)hsums[3 + t]; dpbn[2 + t - 1] = (float)hsums[t]; }
    }
    const int bip = t >> 6, sq = t & 63;
    const int cell = act_cell(sq >> 3, bip, sq & 7);
    for (int pp = 0; pp < ppb; ++pp) {
        const int pair = blockIdx.x * ppb + pp;
        if (pair >= pairs) break;
        const int board = 2 * pair + bip;
        float dr[3] = {0.f, 0.f, 0.f};
        if (board < n) {
            const size_t i = (size_t)board * 64 + sq;
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const float f = h == 0 ? vfeat[i] : pfeat[i * 2 + h - 1];
                const float df = h == 0 ? dvfeat[i] : dpfeat[i * 2 + h - 1];
                const float dz = f > 0.f ? df : df * slope;
                dr[h] = k0[h] * (dz - m1[h] - (head_raw[i * 3 + h] - hmr[h]) * hmr[3 + h] * m2[h]);
            }
        }
        for (int j = 0; j < cj; ++j) {
            const size_t off = cell_off(pair, cj, j, cell);
            float x[8], o[8], prod[32];
            unpack8(__ldg(reinterpret_cast<const uint4*>(a + off)), x);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ch = j * 8 + e;
                o[e] = dr[0] * w[0][ch] + dr[1] * w[1][ch] + dr[2] * w[2][ch];
#pragma unroll
                for (int h = 0; h < 3; ++h) prod[h * 8 + e] = dr[h] * x[e];
                prod[24 + e] = 0.f;
            }
            // the 24 products of this chunk summed over the warp's 32 squares, transposed: lane h * 8 + e ends up with the total
            warp_transpose_sum32_t(prod, lane);
            if (lane < 24) atomicAdd(&dw[lane >> 3][j * 8 + (lane & 7)], prod[0]);
            if (board < n) store_row_cell(gout, off, t, pack8(o));
        }
    }
    __syncthreads();
    for (int i = t; i < c; i += 128) {
        atomicAdd(dwv + i, dw[0][i]);
        atomicAdd(dwp + i * 2, dw[1][i]);
        atomicAdd(dwp + i * 2 + 1, dw[2][i]);
    }
}

// ------------------------------------------------------------------------------------------------ optimizer, weight packing
// SGD(momentum, nesterov=True) of model.py:59 with the l2(4e-4) kernel regularizers folded into the gradient:
// g = grad * gscale + 2 * l2 * w (first n_reg parameters: the kernels); v = mom * v - lr * g; w += nesterov ? mom * v - lr * g : v.
// losses[2] += sum of w^2 over the regularized parameters (before the update).
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ w, float* __restrict__ v, const float* __restrict__ g, size_t n, size_t n_reg,
                                                   const float* __restrict__ hp /* lr, momentum, l2, grad scale, nesterov: device memory, so
                                                   the launch is the same every step and can sit in a CUDA graph */,
                                                   double* __restrict__ losses) {
    __shared__ float red[8];
    const float lr = hp[0], mom = hp[1], l2 = hp[2], gscale = hp[3];
    const bool nesterov = hp[4] != 0.f;
    float sq[1] = {0.f};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float wi = w[i];
        float gi = g[i] * gscale;
        if (i < n_reg) { gi = fmaf(2.f * l2, wi, gi); sq[0] = fmaf(wi, wi, sq[0]); }
        const float vi = mom * v[i] - lr * gi;
        v[i] = vi;
        w[i] = wi + (nesterov ? mom * vi - lr * gi : vi);
    }
    const double s = block_sum<1>(sq, red);
    if (threadIdx.x == 0) atomicAdd(losses + 2, s);
}

// f32 HWIO kernel [9][cin][c] -> the forward pack of the tower kernel (conv_weight_index; stem: bf16 hi + lo halves and the
// remainder channels 119 / 120) and, for the residual layers, the data-gradient pack: taps flipped, channels transposed
__global__ void pack_conv_kernel(const float* __restrict__ w, int cin, int cin_pad, int c, int n, int stem, __nv_bfloat16* __restrict__ fwd,
                                 __nv_bfloat16* __restrict__ tr) {
    const size_t total = (size_t)9 * cin_pad * n;
    const int kc_in = cin_pad / 64;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx % n), ci = (int)((idx / n) % cin_pad), tap = (int)(idx / ((size_t)n * cin_pad));
        int src = ci;
        if (stem) src = ci == 119 ? 113 : (ci == 120 ? 118 : (ci > 120 ? -1 : ci));
        else if (ci >= cin) src = -1;
        const float wf = (src >= 0 && co < c) ? w[((size_t)tap * cin + src) * c + co] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16(wf);
        fwd[conv_weight_index(ci >> 6, tap, ci & 63, co, n)] = hi;
        if (stem) fwd[conv_weight_index((ci >> 6) + kc_in, tap, ci & 63, co, n)] = __float2bfloat16(wf - __bfloat162float(hi));
        if (tr) tr[conv_weight_index(co >> 6, 8 - tap, co & 63, ci, cin_pad)] = hi;
    }
}

__global__ void fill_kernel(float* p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ launchers
static dim3 chunk_grid(int pairs, int cj) {
    int splits = 2368 / cj;  // ~16 blocks of 128 threads per SM
    if (splits > pairs) splits = pairs;
    if (splits < 1) splits = 1;
    return dim3(cj, splits);
}
#define CB_LAUNCH_OK()                 \
    do {                               \
        CB_CUDA_OK(cudaGetLastError()); \
        return 0;                      \
    } while (0)

int launch_planes_from_images(const int16_t* img, int n, __nv_bfloat16* planes, cudaStream_t st) {
    planes_from_images_kernel<<<592, 256, 0, st>>>(img, n, (uint8_t*)planes);
    CB_LAUNCH_OK();
}
int launch_bn_stats(const __nv_bfloat16* y, int pairs, int cj, double* stats, cudaStream_t st) {
    bn_stats_kernel<<<chunk_grid(pairs, cj), SQ_THREADS, 0, st>>>((const uint8_t*)y, pairs, cj, stats);
    CB_LAUNCH_OK();
}
int launch_bn_apply(const __nv_bfloat16* y, const __nv_bfloat16* skip, __nv_bfloat16* out, int n, int cj, int c, const double* stats, const float* gamma,
                    const float* beta, float* mean_rstd, float* moving, float eps, float slope, float momentum, cudaStream_t st) {
    const int pairs = (n + 1) / 2;
    bn_apply_kernel<<<chunk_grid(pairs, cj), SQ_THREADS, 0, st>>>((const uint8_t*)y, (const uint8_t*)skip, (uint8_t*)out, n, pairs, cj, c, stats, gamma, beta,
                                                          mean_rstd, moving, eps, slope, momentum);
    CB_LAUNCH_OK();
}
int launch_bn_bwd_reduce(const __nv_bfloat16* g, const __nv_bfloat16* a, const __nv_bfloat16* y, int n, int cj, const float* mean_rstd, float slope,
                         double* sums, cudaStream_t st) {
    const int pairs = (n + 1) / 2;
    bn_bwd_reduce_kernel<<<chunk_grid(pairs, cj), SQ_THREADS, 0, st>>>((const uint8_t*)g, (const uint8_t*)a, (const uint8_t*)y, n, pairs, cj, mean_rstd, slope, sums);
    CB_LAUNCH_OK();
}
int launch_bn_bwd_apply(const __nv_bfloat16* g, const __nv_bfloat16* a, const __nv_bfloat16* y, __nv_bfloat16* dy, __nv_bfloat16* dz, int n, int cj, int c,
                        const float* mean_rstd, const float* gamma, const double* sums, float slope, float* dgamma, float* dbeta, cudaStream_t st) {
    const int pairs = (n + 1) / 2;
    bn_bwd_apply_kernel<<<chunk_grid(pairs, cj), SQ_THREADS, 0, st>>>((const uint8_t*)g, (const uint8_t*)a, (const uint8_t*)y, (uint8_t*)dy, (uint8_t*)dz, n, pairs,
                                                              cj, c, mean_rstd, gamma, sums, slope, dgamma, dbeta);
    CB_LAUNCH_OK();
}
int launch_head_conv_fwd(const __nv_bfloat16* a, int n, int cj, const float* wv, const float* wp, int c, float* head_raw, cudaStream_t st) {
    head_conv_fwd_kernel<<<(n + 1) / 2, 128, 0, st>>>((const uint8_t*)a, n, cj, wv, wp, c, head_raw);
    CB_LAUNCH_OK();
}
int launch_head_stats(const float* head_raw, int n, double* stats, cudaStream_t st) {
    head_stats_kernel<<<148, 256, 0, st>>>(head_raw, n, stats);
    CB_LAUNCH_OK();
}
int launch_head_features_train(const float* head_raw, int n, const double* stats, const float* vbn, const float* pbn, float* vmoving, float* pmoving,
                               const float* fc1, const float* fc2, int c, float eps, float slope, float momentum, float* hmr, float* vfeat,
                               float* pfeat, float* h1pre, float* value, cudaStream_t st) {
    head_features_train_kernel<<<n, 256, 0, st>>>(head_raw, n, stats, vbn, pbn, vmoving, pmoving, fc1, fc2, c, eps, slope, momentum, hmr, vfeat, pfeat,
                                                  h1pre, value);
    CB_LAUNCH_OK();
}
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t mn, float* __restrict__ C) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int z = 0; z < splits; ++z) acc += part[(size_t)z * mn + i];
        C[i] = acc;
    }
}

// scratch: split-K slabs (scratch_floats capacity); a small output with a long reduction (dpfeat: M = batch, N = 128, K = 4672; dfc1:
// 64 x C, K = batch) would otherwise run on a handful of CTAs
int launch_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int ta, int tb, float* scratch, size_t scratch_floats,
                 cudaStream_t st) {
    if (((ta ? M : K) & 3) || ((tb ? K : N) & 3)) { cb_set_error("gemm: contiguous dimensions must be multiples of 4 (filters % 4 == 0)"); return -1; }
    const size_t big = (size_t)((N + 127) / 128) * ((M + 127) / 128);
    if (big >= 148) {
        gemm_tf32_kernel<128, 128><<<dim3((N + 127) / 128, (M + 127) / 128), 256, 0, st>>>(A, B, C, M, N, K, ta, tb, K);
        CB_LAUNCH_OK();
    }
    const int nb = (N + 63) / 64, mb = (M + 63) / 64;
    int splits = (296 + nb * mb - 1) / (nb * mb);  // about two CTAs per SM
    if (splits > K / 128) splits = K / 128;
    if (scratch && (size_t)splits * M * N > scratch_floats) splits = (int)(scratch_floats / ((size_t)M * N));
    if (!scratch || splits < 2) {
        gemm_tf32_kernel<64, 64><<<dim3(nb, mb), 256, 0, st>>>(A, B, C, M, N, K, ta, tb, K);
        CB_LAUNCH_OK();
    }
    const int k_per = ((K + splits - 1) / splits + 31) / 32 * 32;
    splits = (K + k_per - 1) / k_per;
    gemm_tf32_kernel<64, 64><<<dim3(nb, mb, splits), 256, 0, st>>>(A, B, scratch, M, N, K, ta, tb, k_per);
    CB_CUDA_OK(cudaGetLastError());
    const size_t mn = (size_t)M * N;
    splitk_reduce_kernel<<<(unsigned)((mn + 255) / 256 < 1184 ? (mn + 255) / 256 : 1184), 256, 0, st>>>(scratch, splits, mn, C);
    CB_LAUNCH_OK();
}
int launch_softmax_ce(float* logits, const float* pi, int n, double* losses, cudaStream_t st) {
    softmax_ce_kernel<<<n, 256, 0, st>>>(logits, pi, n, losses);
    CB_LAUNCH_OK();
}
int launch_value_loss_bwd(const float* value, const float* z, const float* h1pre, const float* fc2, int n, int c, float slope, float* dpre, float* dh1pre,
                          float* dfc2, double* losses, cudaStream_t st) {
    value_loss_bwd_kernel<<<n, 256, 0, st>>>(value, z, h1pre, fc2, n, c, slope, dpre, dh1pre, losses);
    CB_CUDA_OK(cudaGetLastError());
    CB_CUDA_OK(cudaMemsetAsync(dfc2, 0, (size_t)c * 4, st));
    value_fc2_grad_kernel<<<dim3((c + 63) / 64, 64), 64, 0, st>>>(h1pre, dpre, n, c, slope, dfc2);
    CB_LAUNCH_OK();
}
int launch_head_bn_bwd_reduce(const float* head_raw, const float* vfeat, const float* pfeat, const float* dvfeat, const float* dpfeat, int n,
                              const float* hmr, float slope, double* hsums, cudaStream_t st) {
    head_bn_bwd_reduce_kernel<<<148, 256, 0, st>>>(head_raw, vfeat, pfeat, dvfeat, dpfeat, n, hmr, slope, hsums);
    CB_LAUNCH_OK();
}
int launch_head_conv_bwd(const __nv_bfloat16* a, __nv_bfloat16* gout, int n, int cj, int c, const float* head_raw, const float* vfeat, const float* pfeat,
                         const float* dvfeat, const float* dpfeat, const float* hmr, const double* hsums, const float* vbn, const float* pbn,
                         const float* wv, const float* wp, float slope, float* dwv, float* dwp, float* dvbn, float* dpbn, cudaStream_t st) {
    const int pairs = (n + 1) / 2, ppb = 4;
    CB_CUDA_OK(cudaMemsetAsync(dwv, 0, (size_t)c * 4, st));
    CB_CUDA_OK(cudaMemsetAsync(dwp, 0, (size_t)c * 8, st));
    head_conv_bwd_kernel<<<(pairs + ppb - 1) / ppb, 128, 0, st>>>((const uint8_t*)a, (uint8_t*)gout, n, pairs, cj, c, ppb, head_raw, vfeat, pfeat, dvfeat,
                                                                 dpfeat, hmr, hsums, vbn, pbn, wv, wp, slope, dwv, dwp, dvbn, dpbn);
    CB_LAUNCH_OK();
}
int launch_sgd(float* w, float* v, const float* g, size_t n, size_t n_reg, const float* hp_dev, double* losses, cudaStream_t st) {
    sgd_kernel<<<592, 256, 0, st>>>(w, v, g, n, n_reg, hp_dev, losses);
    CB_LAUNCH_OK();
}
int launch_pack_conv(const float* w, int cin, int cin_pad, int c, int n, int stem, __nv_bfloat16* fwd, __nv_bfloat16* tr, cudaStream_t st) {
    pack_conv_kernel<<<296, 256, 0, st>>>(w, cin, cin_pad, c, n, stem, fwd, tr);
    CB_LAUNCH_OK();
}
int launch_fill(float* p, size_t n, float v, cudaStream_t st) {
    fill_kernel<<<148, 256, 0, st>>>(p, n, v);
    CB_LAUNCH_OK();
}

}  // namespace cb
