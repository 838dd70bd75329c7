// K2 — the residual tower of model.py:19-28,35-37 (3x3 "same" convolution + BatchNorm + (residual add) +
// LeakyReLU, layer after layer) as ONE persistent kernel of implicit GEMMs on tcgen05 tensor cores.
//
//   D[m, co] = sum_{tap, ci} A_tap[m, ci] * W[tap][ci][co]      m = (row, board, col) of 2 boards = 128 rows
//
// * A is never materialised: the activations of a board pair sit in HBM in the tile-native layout
//   (common.cuh) with a zero border, one TMA tile load brings 64 channels of the pair into shared memory
//   byte for byte, and each of the nine taps is just a different start address of the K-major /
//   no-swizzle UMMA descriptor (LBO 3200 B between the two 8-channel core matrices, SBO 160 B between
//   8-pixel groups).
// * Clusters of two CTAs (one TPC) issue cta_group::2 MMAs of M = 256: each CTA owns one board pair (128
//   accumulator rows in its own TMEM) and stages only HALF of every weight block (N/2 output channels,
//   pre-packed per (k-chunk, tap, half) as [8][N/2][8] bf16); the TMA loads of both CTAs complete on the
//   leader's barriers, the leader's single elected lane issues the MMAs, tcgen05.commit multicasts the
//   "slot free" / "accumulator ready" signals back to both CTAs.
// * TMEM holds TWO accumulator stages of 256 columns: while the epilogue drains stage s, the MMAs of the
//   next work item fill stage s^1, so BN / skip / LeakyReLU / stores overlap the tensor pipe.
// * A 3x3 convolution never mixes boards, so a board pair can run through ALL layers without any
//   grid-wide synchronisation.  The (unit = 4 boards, layer) items are cut into equal contiguous ranges
//   per cluster (perfect balance over the whole tower instead of a ragged last wave per layer); a
//   cluster walks its range layer by layer over its units, and a unit shared between two neighbouring
//   clusters is handed over through a per-board-pair progress flag in global memory.  The same flag
//   orders "epilogue wrote layer l" before "TMA reads layer l as the input of layer l+1".
// * Warp roles per CTA: warp 0 = TMA producer of the activations, warp 2 = TMA producer of the weights (independent of the
//   progress flags), warp 3 = publisher of the progress flags, warp 1 = MMA issuer (leader CTA), warps 4..11 = epilogue
//   (tcgen05.ld -> BN scale/shift -> +skip -> LeakyReLU -> bf16 -> HBM in the same tile-native layout; on
//   the last layer also the three 1x1 head convolutions of model.py:43,52 as dot products over the fp32
//   row).  Two epilogue warps share a TMEM lane quarter and split the channel batches.  The four control warps are one
//   warpgroup and hand most of their registers to the two epilogue warpgroups (setmaxnreg), which hold the prefetched
//   skip tiles of a whole item in registers.
// * The residual stream x_{i+1} = LReLU(x_i + branch) crosses all blocks: rounding it to bf16 at every block is the largest
//   error term of a bf16 tower (profiles/r2_precision_variants.json: max |dvalue| 3.3e-2 -> 1.2e-2 at 20x256).  So the stream
//   is kept as bf16 hi (the tile the next convolution's TMA reads) + bf16 lo = bf16(x - hi) in a second buffer of the same
//   layout that only the epilogue touches, in place (TowerParams::lo_buf): the skip add sees ~16 mantissa bits.
#include <stdlib.h>

#include "kernels.cuh"

namespace cb {

constexpr int CONV_THREADS = 384;  // activation producer, MMA issuer, weight producer, flag publisher, 8 epilogue warps
constexpr int CTRL_REGS = 56, EPI_REGS = 224;  // setmaxnreg split: 128 * 56 + 256 * 224 <= 64 K registers
constexpr int PUB_SLOTS = 8;
constexpr int A_SLOTS = 2, B_SLOTS = 16;  // weight ring: up to 16 stages inside B_RING_BYTES
constexpr int A_SLOT_BYTES = ACT_K64_BYTES;  // one board pair x 64 channels
constexpr int B_RING_BYTES = 9 * 128 * 128;  // stages of three taps: N = 256: 3 x 48 KB per CTA, N = 128: 6 x 24 KB, N = 64: 12 x 12 KB
constexpr int EPI_WARPS = 8;

struct ConvSmem {
    uint64_t a_full[A_SLOTS], a_empty[A_SLOTS];
    uint64_t b_full[B_SLOTS], b_empty[B_SLOTS];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t pub_full[PUB_SLOTS], pub_empty[PUB_SLOTS];  // epilogue warps -> flag publisher
    uint32_t tmem_base;
};
static_assert(sizeof(ConvSmem) <= 1024, "barrier block");
constexpr int CONV_SMEM_PARAM_FLOATS = 256 * 3 + 2 * 128 * 3 + EPI_WARPS * 512;  // head weights, head partials, per-warp BN scale / shift of two layers
constexpr size_t CONV_SMEM_BYTES = 1024 + A_SLOTS * A_SLOT_BYTES + B_RING_BYTES + CONV_SMEM_PARAM_FLOATS * 4;

// This cluster's share of the flattened (unit, layer) items i = unit * nl + layer: the range [lo, hi).
struct WorkRange {
    int lo, hi, u_first, nu, nl;
    __device__ WorkRange(int cid, int nclusters, int units, int nl_) : nl(nl_) {
        const long long total = (long long)units * nl_;
        lo = (int)(total * cid / nclusters);
        hi = (int)(total * (cid + 1) / nclusters);
        u_first = lo / nl_;
        nu = (hi - 1) / nl_ - u_first + 1;
    }
    // k-th unit of a layer round: the unit the NEXT cluster continues goes first (its flag is published early),
    // the unit the PREVIOUS cluster began goes last (its flag has had a whole round to arrive)
    __device__ int unit(int k) const { return nu == 1 ? u_first : (k == 0 ? u_first + nu - 1 : (k == nu - 1 ? u_first : u_first + k)); }
    __device__ bool mine(int u, int l) const { const int i = u * nl + l; return i >= lo && i < hi; }
    // The order every role walks this cluster's items in: rounds r = 0 .. nl, unit slots k = 0 .. nu - 1 within a round, and the odd
    // slots run ONE LAYER BEHIND the even ones (layer = r - (k & 1)).  Consecutive items then alternate between a layer with a
    // residual add (long epilogue: skip tile read, hi + lo written) and one without (short epilogue), so the two TMEM accumulator
    // stages average the epilogue time out; layer-major order made 3-4 long epilogues in a row, each longer than its MMAs.
    // (u, l) still follows (u, l - 1) by a whole round.
    __device__ bool item(int r, int k, int& u, int& l) const {
        l = r - (k & 1);
        if (l < 0 || l >= nl) return false;
        u = unit(k);
        return mine(u, l);
    }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// sum over the 32 lanes of a warp of 32 values per lane, transposed: afterwards lane c holds the total of value c in v[0].
// Butterfly with halving payload: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ void warp_transpose_sum32(float (&v)[32], int lane) {
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

// One k-chunk (64 input channels) of one item: the 9 / TPS weight stages of TPS taps each, four K = 16 MMAs per tap.  Unrolled: the tap's
// start cell (1 + dy) * 20 + 1 + dx inside the activation tile and the K steps are immediates added to the descriptors.
template <int TPS, bool FP8>
__device__ __forceinline__ void mma_k_chunk(ConvSmem* s, uint32_t d, uint64_t a_desc, uint64_t b_desc0, uint32_t idesc, uint32_t slot16, uint32_t tap16,
                                            uint32_t b_ks, int nbs, uint32_t& bs, uint32_t& bph, uint32_t accumulate) {
    constexpr uint32_t a_ks = 2 * ACT_CHUNK_BYTES / 16;
#pragma unroll
    for (int sg = 0; sg < 9 / TPS; ++sg) {
        ptx::mbar_wait(&s->b_full[bs], bph);
        ptx::tc_fence_after();
        const uint64_t bd_slot = b_desc0 + (uint64_t)(bs * slot16);
        if (ptx::elect_one()) {
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
                const int tap = sg * TPS + t;
                const uint64_t ad0 = a_desc + (uint64_t)((tap / 3) * 20 + (tap % 3));
                const uint64_t bd0 = bd_slot + (uint64_t)(t * tap16);  // one tap = N/2 rows x 128 B
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    if constexpr (FP8) ptx::umma2_e4m3(d, ad0 + ks * a_ks, bd0 + ks * b_ks, idesc, (tap | ks) != 0 ? 1u : accumulate);
                    else ptx::umma2_bf16(d, ad0 + ks * a_ks, bd0 + ks * b_ks, idesc, (tap | ks) != 0 ? 1u : accumulate);
            }
            ptx::umma2_commit_mc(&s->b_empty[bs]);  // weight stage of BOTH CTAs is free when these MMAs retire
        }
        __syncwarp();
        if (++bs == (uint32_t)nbs) { bs = 0; bph ^= 1; }
    }
}

// STATS: the training forward pass (one layer per launch, identity epilogue, no residual add): the epilogue also accumulates the
// per-channel sum and sum of squares of its bf16-rounded outputs (BatchNormalization's batch statistics), so the raw convolution
// is not read again for them.
template <bool STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1) tower_tcgen05_kernel(const __grid_constant__ TowerParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    ConvSmem* s = reinterpret_cast<ConvSmem*>(smem);
    uint8_t* a_buf = smem + 1024;
    uint8_t* b_buf = a_buf + A_SLOTS * A_SLOT_BYTES;
    float* s_head = reinterpret_cast<float*>(b_buf + B_RING_BYTES);
    float* s_hpart = s_head + 768;  // [2 stages][128 rows][3]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    // the search loop passes the number of boards of this step on the device (the batch shrinks as trees finish): no host round trip
    const int pairs = p.n_boards_dev ? min(p.pairs, (*p.n_boards_dev + 1) >> 1) : p.pairs;
    const int units = (pairs + 1) / 2;  // one unit = two board pairs = one M=256 accumulator stage
    const int n = p.n, nh = n >> 1, nl = p.nl;
    // a weight stage = three taps of one k-chunk, fetched as `boxes` TMA boxes of `tpb` taps (kernels.cuh: conv_taps_per_stage)
    constexpr int tps = 3, stages = 3;
    const int tpb = conv_taps_per_box(n), boxes = tps / tpb;
    const uint32_t slot_bytes = (uint32_t)tps * nh * 128;
    const int nbs = min(B_SLOTS, B_RING_BYTES / (int)slot_bytes);
    const WorkRange w(cid, nclusters, units, nl);
    const int flag_base = p.epoch << 8;
    // throughput probe only (TowerParams::fp8_probe): an e4m3 layer has half the k-chunks (128 channels per 128-byte row)
    const int kc_shift = (!STATS && (p.fp8_probe & 1)) ? 1 : 0;

    if (STATS)
        for (int i = threadIdx.x; i < 512; i += CONV_THREADS) s_head[i] = 0.f;  // [2][256] running sums of this CTA (head weights are not used when training)
    if (!STATS && p.head_w)
        for (int i = threadIdx.x; i < n; i += CONV_THREADS) {
            s_head[i] = p.head_w[i];
            s_head[256 + i] = p.head_w[n + i];
            s_head[512 + i] = p.head_w[2 * n + i];
        }
    if (threadIdx.x == 0) {
        for (int i = 0; i < A_SLOTS; ++i) { ptx::mbar_init(&s->a_full[i], 1); ptx::mbar_init(&s->a_empty[i], 1); }
        for (int i = 0; i < B_SLOTS; ++i) { ptx::mbar_init(&s->b_full[i], 1); ptx::mbar_init(&s->b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s->acc_full[i], 1);
            ptx::mbar_init(&s->acc_empty[i], 2 * EPI_WARPS);  // every epilogue warp of both CTAs
        }
        for (int i = 0; i < PUB_SLOTS; ++i) { ptx::mbar_init(&s->pub_full[i], EPI_WARPS); ptx::mbar_init(&s->pub_empty[i], 1); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc2(&s->tmem_base, 512);
        ptx::tmem_relinquish2();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
    ptx::tc_fence_after();
    const uint32_t tmem = s->tmem_base;

    if (warp < 4) {
    // the inference epilogue keeps hi + lo skip tiles of a whole item in registers: the control warpgroup gives its own away
    if constexpr (!STATS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CTRL_REGS));
    if (warp == 0) {
        // ------------------------------------------------------------------ activation producer (both CTAs; whole warp, one elected lane issues)
        uint32_t as = 0, aph = 1;
        // operands of BOTH CTAs complete on the leader's "full" barriers (the MMA issuer lives there)
        const uint32_t a_full0 = ptx::mapa(ptx::smem_u32(&s->a_full[0]), 0);
        // flat walk over this cluster's items with one item of look-ahead: the progress flag of the NEXT item is read
        // (ld.acquire.gpu, an L2 round trip) while this item's loads are being issued
        int r = 0, k = -1, u = 0, l = 0;
        auto advance = [&](int& rr, int& kk, int& uu, int& ll) {  // next item of this cluster, rr > nl at the end
            do {
                if (++kk == w.nu) { kk = 0; ++rr; }
            } while (rr <= nl && !w.item(rr, kk, uu, ll));
        };
        advance(r, k, u, l);
        int seen = r <= nl && l > 0 ? ld_acquire_gpu(p.flags + 2 * u + rank) : 0;
        int l_loaded = -1;
        LayerDesc L{};
        while (r <= nl) {
            if (l != l_loaded) { L = p.layers[l]; l_loaded = l; }
            const CUtensorMap* tm_in = &p.tm_act[L.in_buf];
            const int pair = min(2 * u + (int)rank, pairs - 1);  // odd tail: the peer re-reads the last pair, its result is dropped
            if (l > 0) {
                // the input of this item is the output of (unit, layer - 1): written by this CTA's epilogue warps, or by
                // the same-rank CTA of the previous cluster; either way its progress flag says when it is in L2
                const int* flag = p.flags + 2 * u + rank;
                while (seen < flag_base + l) seen = ld_acquire_gpu(flag);
                fence_proxy_async();
            }
            int nr = r, nk_ = k, nu_ = u, nl_ = l;
            advance(nr, nk_, nu_, nl_);
            if (nr <= nl && nl_ > 0) seen = ld_acquire_gpu(p.flags + 2 * nu_ + rank);  // consumed at the top of the next item
            for (int kc = 0; kc < max(1, L.kc >> kc_shift); ++kc) {
                ptx::mbar_wait_parked(&s->a_empty[as], aph);
                if (ptx::elect_one()) {
                    if (rank == 0) ptx::mbar_expect_tx(&s->a_full[as], 2 * ACT_K64_BYTES);
                    ptx::tma_load_2d_pair(a_buf + as * A_SLOT_BYTES, tm_in, 0, (pair * L.cj_in + 8 * (kc % L.kc_in)) * (ACT_CHUNK_BYTES / 128),
                                          a_full0 + as * 8);
                }
                if (++as == A_SLOTS) { as = 0; aph ^= 1; }
            }
            r = nr; k = nk_; u = nu_; l = nl_;
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ weight producer (both CTAs): the weight stream does not depend on
        // any activation, so it runs ahead of the progress flags the activation producer has to wait for
        uint32_t bs = 0, bph = 1;  // ring slot and the parity of its "empty" barrier (a fresh barrier passes a wait on parity 1)
        const uint32_t b_full0 = ptx::mapa(ptx::smem_u32(&s->b_full[0]), 0);
        for (int r = 0; r <= nl; ++r) {
            for (int k = 0; k < w.nu; ++k) {
                int u, l;
                if (!w.item(r, k, u, l)) continue;
                const int kcs = max(1, p.layers[l].kc >> kc_shift);
                const CUtensorMap* tm_w = &p.maps[p.layers[l].w_map];
                for (int kc = 0; kc < kcs; ++kc) {
                    const int wrow = (kc * 2 + (int)rank) * 9 * nh;  // [kc][half][tap][8][N/2][8]
                    for (int sg = 0; sg < stages; ++sg) {
                        ptx::mbar_wait_parked(&s->b_empty[bs], bph);
                        if (ptx::elect_one()) {
                            if (rank == 0) ptx::mbar_expect_tx(&s->b_full[bs], 2 * slot_bytes);
                            for (int bx = 0; bx < boxes; ++bx)
                                ptx::tma_load_2d_pair(b_buf + bs * slot_bytes + bx * tpb * nh * 128, tm_w, 0, wrow + (sg * tps + bx * tpb) * nh, b_full0 + bs * 8);
                        }
                        if (++bs == (uint32_t)nbs) { bs = 0; bph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // -------------------------------------------------------------- MMA issuer (leader CTA; whole warp, one elected lane issues)
            // The loop that feeds the tensor pipe must itself stay well under the 512 tensor cycles a weight stage of four N = 256 MMAs
            // lasts (ncu: with runtime `% nbs` ring arithmetic it took ~70 dependent instructions per one-tap stage and the issuing warp was
            // busy most of the time): ring slots and phases are counters, the nine taps are unrolled so that their descriptor
            // offsets are immediates.
            const bool probe = !STATS && (p.fp8_probe & 1);
            const uint32_t idesc = probe ? ptx::idesc_e4m3_f32(256, n) : ptx::idesc_bf16_f32(256, n);
            // K-major / no-swizzle descriptors: hi word = SBO | version, lo word = address | LBO; a K=16 step adds to the address field only
            const uint64_t a_desc0 = ptx::smem_desc(ptx::smem_u32(a_buf), ACT_CHUNK_BYTES, 160);
            const uint64_t b_desc0 = ptx::smem_desc(ptx::smem_u32(b_buf), (uint32_t)nh * 16, 128);
            const uint32_t b_ks = 2 * nh;  // descriptor address units of 16 B
            const uint32_t slot16 = slot_bytes / 16, tap16 = (uint32_t)nh * 8;
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, j = 0;
            for (int r = 0; r <= nl; ++r) {
                for (int k = 0; k < w.nu; ++k) {
                    int u, l;
                    if (!w.item(r, k, u, l)) continue;
                    const int kcs = max(1, p.layers[l].kc >> kc_shift);
                    const uint32_t st = j & 1;
                    ptx::mbar_wait(&s->acc_empty[st], ((j >> 1) & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d = tmem + st * 256;
                    for (int kc = 0; kc < kcs; ++kc) {
                        ptx::mbar_wait(&s->a_full[as], aph);
                        const uint64_t a_desc = a_desc0 + (uint64_t)(as * (A_SLOT_BYTES / 16));
                        if (probe) mma_k_chunk<tps, true>(s, d, a_desc, b_desc0, idesc, slot16, tap16, b_ks, nbs, bs, bph, kc != 0);
                        else mma_k_chunk<tps, false>(s, d, a_desc, b_desc0, idesc, slot16, tap16, b_ks, nbs, bs, bph, kc != 0);
                        if (ptx::elect_one()) ptx::umma2_commit_mc(&s->a_empty[as]);
                        __syncwarp();
                        if (++as == A_SLOTS) { as = 0; aph ^= 1; }
                    }
                    if (ptx::elect_one()) ptx::umma2_commit_mc(&s->acc_full[st]);
                    __syncwarp();
                    ++j;
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ flag publisher (both CTAs): when all eight epilogue warps have
        // stored item j, release-store the progress flag of its board pair (the release is cumulative over the mbarrier)
        // for the activation producer of (unit, l + 1), here or in the next cluster
        if (lane == 0) {
            uint32_t j = 0;
            for (int r = 0; r <= nl; ++r)
                for (int k = 0; k < w.nu; ++k) {
                    int u, l;
                    if (!w.item(r, k, u, l)) continue;
                    ptx::mbar_wait_parked(&s->pub_full[j % PUB_SLOTS], (j / PUB_SLOTS) & 1);
                    ptx::mbar_arrive(&s->pub_empty[j % PUB_SLOTS]);
                    fence_proxy_async();  // the epilogue's generic stores (observed through the mbarrier) before the TMA reads that follow the flag
                    st_release_gpu(p.flags + 2 * u + rank, flag_base + l + 1);
                    ++j;
                }
        }
    }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 4..11, both CTAs)
        if constexpr (!STATS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
        // Two warps per TMEM lane quarter (a warp may only read lanes 32*(warp%4)..+31) split the 32-channel
        // batches of the row: warp half h takes batches h, h+2, ...  The skip tile is read through a 4-deep
        // register ring whose first loads are issued BEFORE the accumulator is ready.
        const int q = warp & 3;          // TMEM lane quarter this warp may read
        const int h = (warp - 4) >> 2;   // which half of the channel batches
        const int m = q * 32 + lane;     // accumulator row
        const int g = m >> 3, col = m & 7, row = g >> 1, bip = g & 1;
        const int cell = act_cell(row, bip, col);
        const int edge = col == 0 ? -16 : (col == 7 ? 16 : 0);  // byte offset of the zero border cell beside this lane's, if any (one store for both)
        const int cjo = n >> 3, ncb = n >> 5;
        const int nk = (ncb - h + 1) >> 1;  // batches of this warp: cb = h + 2k
        const uint32_t acc_empty0 = ptx::mapa(ptx::smem_u32(&s->acc_empty[0]), 0);
        // folded BN scale / shift of THIS warp's channel batches, staged in a warp-private slice of shared memory once per layer: read
        // from global memory in the batch loop they cost an L1 / L2 round trip in front of every group of FFMAs (ncu: half of the
        // epilogue's stall samples), and the epilogue is what bounds the layers with a residual add
        // two slices, one per layer parity: consecutive items alternate between two layers (WorkRange::item)
        float* s_bn2 = s_hpart + 768 + (warp - 4) * 512;  // per slice: [0..127] scale, [128..255] shift of batches h, h + 2, ...
        int bn_l0 = -1, bn_l1 = -1;  // layer whose scale / shift each slice holds
        uint32_t j = 0;
        for (int rd = 0; rd <= nl; ++rd)
        for (int k = 0; k < w.nu; ++k) {
            int u, l;
            if (!w.item(rd, k, u, l)) continue;
            const LayerDesc L = p.layers[l];
            float* s_bn = s_bn2 + (l & 1) * 256;
            if (((l & 1) ? bn_l1 : bn_l0) != l) {
                __syncwarp();
                for (int kb = 0; kb < nk; ++kb) {
                    s_bn[kb * 32 + lane] = __ldg(L.scale + (h + 2 * kb) * 32 + lane);
                    s_bn[128 + kb * 32 + lane] = __ldg(L.shift + (h + 2 * kb) * 32 + lane);
                }
                __syncwarp();
                if (l & 1) bn_l1 = l; else bn_l0 = l;
            }
            const uint8_t* skip_buf = (!STATS && L.skip_buf >= 0) ? p.bufs[L.skip_buf] : nullptr;
            uint8_t* out_buf = p.bufs[L.out_buf];
            const bool has_skip = !STATS && skip_buf != nullptr;
            // residual stream in two bf16 halves: layers whose output IS the stream (stem, second conv of a block) also store
            // lo = bf16(x - hi); the skip add of the next block reads hi + lo.  Same thread, same cell: updated in place.
            uint8_t* lo_buf = STATS ? nullptr : p.lo_buf;
            // p.lo_from_layer: the stream carries a low half from the output of that layer on (0 = from the stem)
            const bool skip_lo = has_skip && lo_buf != nullptr && l - 2 >= p.lo_from_layer && !(p.fp8_probe & 4);
            const bool write_lo = lo_buf != nullptr && (has_skip || l == 0) && !(nl == 1) && l >= p.lo_from_layer && !(p.fp8_probe & 2);
            const bool heads = !STATS && p.head_w != nullptr && l == nl - 1;
            {
                const uint32_t st = j & 1;
                const int pair = 2 * u + (int)rank;
                const bool active = pair < pairs;
                // the skip tile is the output of (unit, layer - 2): if another cluster wrote it, it is only known to be
                // complete once this item's own operands have been loaded (the producer waited for the flag)
                const bool skip_early = has_skip && w.mine(u, l - 2);
                uint4 sk[4][4], sl[4][4];
                auto load_skip = [&](int r, int cb) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const size_t off = act_offset_bytes(pair, cjo, cb * 4 + jj, cell);
                        sk[r][jj] = __ldcg(reinterpret_cast<const uint4*>(skip_buf + off));
                        if (skip_lo) sl[r][jj] = __ldcg(reinterpret_cast<const uint4*>(lo_buf + off));
                    }
                };
                if (active && skip_early) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (r < nk) load_skip(r, h + 2 * r);
                }
                ptx::mbar_wait_parked(&s->acc_full[st], (j >> 1) & 1);
                ptx::tc_fence_after();
                if (active && has_skip && !skip_early) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (r < nk) load_skip(r, h + 2 * r);
                }
                float h0 = 0.f, h1 = 0.f, h2 = 0.f;
                if (active) {
                    for (int k0 = 0; k0 < nk; k0 += 4) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int kk = k0 + r;
                            if (kk >= nk) break;
                            const int cb = h + 2 * kk;
                            uint32_t acc[32];
                            ptx::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + st * 256 + cb * 32, acc);
                            const float4* sc4 = reinterpret_cast<const float4*>(s_bn + kk * 32);
                            const float4* sh4 = reinterpret_cast<const float4*>(s_bn + 128 + kk * 32);
                            ptx::tmem_wait_ld();
                            float rsum[32], rsq[32];  // STATS only
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                float v[8];
                                const uint32_t* skw = reinterpret_cast<const uint32_t*>(&sk[r][jj]);
                                const uint32_t* slw = reinterpret_cast<const uint32_t*>(&sl[r][jj]);
                                const float4 sa = sc4[jj * 2], sb = sc4[jj * 2 + 1];
                                const float4 ha = sh4[jj * 2], hb = sh4[jj * 2 + 1];
                                const float scl[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                                const float shf[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const int ch = cb * 32 + jj * 8 + e;
                                    float x = fmaf(__uint_as_float(acc[jj * 8 + e]), scl[e], shf[e]);
                                    if (has_skip) {
                                        const uint32_t wv = skw[e >> 1];
                                        float sv = __uint_as_float((e & 1) ? (wv & 0xFFFF0000u) : (wv << 16));
                                        if (skip_lo) {
                                            const uint32_t lv = slw[e >> 1];
                                            sv += __uint_as_float((e & 1) ? (lv & 0xFFFF0000u) : (lv << 16));
                                        }
                                        x += sv;
                                    }
                                    x = x > 0.f ? x : x * p.slope;
                                    v[e] = x;
                                    if (heads) {
                                        h0 = fmaf(x, s_head[ch], h0);
                                        h1 = fmaf(x, s_head[256 + ch], h1);
                                        h2 = fmaf(x, s_head[512 + ch], h2);
                                    }
                                }
                                uint4 o;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
                                o.x = *reinterpret_cast<uint32_t*>(&t0);
                                o.y = *reinterpret_cast<uint32_t*>(&t1);
                                o.z = *reinterpret_cast<uint32_t*>(&t2);
                                o.w = *reinterpret_cast<uint32_t*>(&t3);
                                const size_t ooff = act_offset_bytes(pair, cjo, cb * 4 + jj, cell);
                                uint8_t* dst = out_buf + ooff;
                                *reinterpret_cast<uint4*>(dst) = o;
                                if (write_lo) {
                                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
                                    float d[8];
#pragma unroll
                                    for (int e = 0; e < 8; ++e)
                                        d[e] = v[e] - __uint_as_float((e & 1) ? (ow[e >> 1] & 0xFFFF0000u) : (ow[e >> 1] << 16));
                                    uint4 ol;
                                    __nv_bfloat162 u0 = __floats2bfloat162_rn(d[0], d[1]), u1 = __floats2bfloat162_rn(d[2], d[3]);
                                    __nv_bfloat162 u2 = __floats2bfloat162_rn(d[4], d[5]), u3 = __floats2bfloat162_rn(d[6], d[7]);
                                    ol.x = *reinterpret_cast<uint32_t*>(&u0);
                                    ol.y = *reinterpret_cast<uint32_t*>(&u1);
                                    ol.z = *reinterpret_cast<uint32_t*>(&u2);
                                    ol.w = *reinterpret_cast<uint32_t*>(&u3);
                                    uint8_t* dl = lo_buf + ooff;
                                    *reinterpret_cast<uint4*>(dl) = ol;
                                    // no whole-sector border store here: on skip layers this very thread has just READ the cell's sector (its
                                    // lo input), so the sector is complete in L2 when the store lands; only the stem writes lo blind
                                    if (p.full_sectors && edge && l == 0) *reinterpret_cast<uint4*>(dl + edge) = make_uint4(0u, 0u, 0u, 0u);
                                }
                                // a board row (8 cells, 128 B) starts 16 B into a 32-B sector: the first / last column also store the zero
                                // border cell next to theirs, so every row is five WHOLE sectors and L2 never has to fetch a partially
                                // written sector from DRAM before writing it back (only where the tensors overflow L2: TowerParams::full_sectors)
                                if (p.full_sectors && edge) *reinterpret_cast<uint4*>(dst + edge) = make_uint4(0u, 0u, 0u, 0u);
                                if constexpr (STATS) {  // statistics of what was stored (bf16), as the BN pass will read it back
                                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const float r = __uint_as_float((e & 1) ? (ow[e >> 1] & 0xFFFF0000u) : (ow[e >> 1] << 16));
                                        rsum[jj * 8 + e] = r;
                                        rsq[jj * 8 + e] = r * r;
                                    }
                                }
                            }
                            if constexpr (STATS) {
                                warp_transpose_sum32(rsum, lane);
                                warp_transpose_sum32(rsq, lane);
                                atomicAdd(&s_head[cb * 32 + lane], rsum[0]);        // 4 lane quarters x item after item: little contention
                                atomicAdd(&s_head[256 + cb * 32 + lane], rsq[0]);
                            }
                            if (has_skip && kk + 4 < nk) load_skip(r, h + 2 * (kk + 4));
                        }
                    }
                }
                // the accumulator stage is drained: hand it back to the MMA issuer
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_remote(acc_empty0 + st * 8);
                if (heads) {
                    // the two warps of a lane quarter hold partial dot products over their channel batches
                    float* hp = s_hpart + ((size_t)st * 128 + m) * 3;
                    if (h == 1) { hp[0] = h0; hp[1] = h1; hp[2] = h2; }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                    if (h == 0 && active) {
                        float* ho = p.head_out + ((size_t)(pair * 2 + bip) * 64 + row * 8 + col) * 3;
                        ho[0] = h0 + hp[0]; ho[1] = h1 + hp[1]; ho[2] = h2 + hp[2];
                    }
                }
                // "layer l of this board pair is in memory": the warp reports to the flag publisher (warp 3), which issues
                // the generic->async proxy fence and the device-scope release.  Both cost microseconds; on an epilogue warp
                // they would bound the item rate of narrow layers (measured: 2,000 cycles of accumulator hand-back wait
                // per item at N = 64).
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_wait_parked(&s->pub_empty[j % PUB_SLOTS], ((j / PUB_SLOTS) & 1) ^ 1);
                    ptx::mbar_arrive(&s->pub_full[j % PUB_SLOTS]);
                }
                __syncwarp();
                ++j;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (STATS)  // this CTA's sums over all its items -> the layer's batch statistics
        for (int i = threadIdx.x; i < 2 * n; i += CONV_THREADS) atomicAdd(p.stats + i, (double)s_head[(i >= n ? 256 - n : 0) + i]);
    ptx::cluster_sync();  // the peer may still be reading this CTA's shared memory / signalling its barriers
    if (warp == 1) ptx::tmem_dealloc2(tmem, 512);
}

// Plain CUDA-core restatement of the same layer on the same bf16 inputs (fp32 accumulate).  Test
// infrastructure for the tcgen05 kernel (exported as cb_debug_conv_reference); never on the product path.
__global__ void conv3x3_reference_kernel(ConvParams p) {
    const int boards = p.pairs * 2;
    const size_t total = (size_t)boards * 64 * p.n;
    const int cjo = p.n >> 3;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx % p.n);
        const int sq = (int)((idx / p.n) % 64);
        const int board = (int)(idx / ((size_t)p.n * 64));
        const int pair = board >> 1, bip = board & 1, row = sq >> 3, col = sq & 7;
        float acc = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int cell = act_cell(row + dy, bip, col + dx);  // border cells are zero
            for (int ci = 0; ci < p.kc * 64; ++ci) {
                const int cin = ci % (p.kc_in * 64);
                const __nv_bfloat16 a = *reinterpret_cast<const __nv_bfloat16*>(
                    reinterpret_cast<const uint8_t*>(p.in) + act_offset_bytes(pair, p.cj_in, cin >> 3, cell) + (cin & 7) * 2);
                const int kc = ci >> 6, k = ci & 63;
                const __nv_bfloat16 w = p.wgt[conv_weight_index(kc, tap, k, co, p.n)];
                acc = fmaf(__bfloat162float(a), __bfloat162float(w), acc);
            }
        }
        float x = fmaf(acc, p.scale[co], p.shift[co]);
        const size_t off = act_offset_bytes(pair, cjo, co >> 3, act_cell(row, bip, col)) + (co & 7) * 2;
        if (p.skip) x += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const uint8_t*>(p.skip) + off));
        x = x > 0.f ? x : x * p.slope;
        *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(p.out) + off) = __float2bfloat16_rn(x);
    }
}

// 2-D tensor map over a buffer seen as `rows` rows of 128 B (64 bf16), box = box_rows whole rows, no swizzle:
// the tile lands in shared memory byte for byte as it lies in HBM.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_row_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t box_rows) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { cb_set_error("cuTensorMapEncodeTiled not available"); return -1; }
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[2] = {64, rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { cb_set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return -1; }
    return 0;
}

int conv_make_row_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t box_rows) { return make_row_map(tm, base, rows, box_rows); }

int launch_tower(const TowerParams& p, int num_sms, cudaStream_t stream) {
    static std::atomic<uint64_t> smem_set[2];
    if (cb_opt_in_dynamic_smem((const void*)tower_tcgen05_kernel<false>, (int)CONV_SMEM_BYTES, smem_set[0]) ||
        cb_opt_in_dynamic_smem((const void*)tower_tcgen05_kernel<true>, (int)CONV_SMEM_BYTES, smem_set[1]))
        return -1;
    if (p.n % 64 != 0 || p.n > 256 || p.n < 64 || p.nl < 1 || p.nl > 255 || p.pairs < 1) {
        cb_set_error("tower: unsupported shape");
        return -1;
    }
    const int units = (p.pairs + 1) / 2;
    const int max_clusters = num_sms / 2;
    const int clusters = units < max_clusters ? units : max_clusters;
    if (p.stats) {
        if (p.nl != 1 || p.head_w) { cb_set_error("tower: batch statistics are gathered one layer per launch, without heads"); return -1; }
        tower_tcgen05_kernel<true><<<2 * clusters, CONV_THREADS, CONV_SMEM_BYTES, stream>>>(p);
    } else {
        tower_tcgen05_kernel<false><<<2 * clusters, CONV_THREADS, CONV_SMEM_BYTES, stream>>>(p);
    }
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_conv3x3_reference(const ConvParams& p, cudaStream_t stream) {
    conv3x3_reference_kernel<<<592, 256, 0, stream>>>(p);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cb
