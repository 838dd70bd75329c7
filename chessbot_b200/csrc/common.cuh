// Shared device helpers: PTX wrappers for mbarrier / bulk async copy / tcgen05 (sm_100a), the
// tile-native activation layout, and error plumbing for the C ABI.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

void cb_set_error(const std::string& s);
// Device memory of the library (network weights / activations, node and position pools, trainer buffers) comes from the allocator the
// host installed with cb_set_device_allocator (PyTorch's caching allocator in the Python host), else from cudaMalloc.
int cb_device_alloc(void** ptr, size_t bytes);
void cb_device_free(void* ptr);

#define CB_CUDA_OK(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            cb_set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                 \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: opt in once per (kernel, device) — a process
// may hold contexts on several GPUs (engine.Engine.get(device)) and launch from several actor threads.  `done` is the kernel's own
// bit mask of devices already set (a lost race only repeats the idempotent call).
inline int cb_opt_in_dynamic_smem(const void* kernel, int bytes, std::atomic<uint64_t>& done) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cb_set_error("cudaGetDevice failed"); return -1; }
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { cb_set_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): ") + cudaGetErrorString(e)); return -1; }
    done.fetch_or(bit, std::memory_order_release);
    return 0;
}

// Every entry point runs on its context's device and leaves the caller's current device as it found it.
struct cb_device_guard {
    int prev = -1, dev = -1;
    bool ok = true;
    explicit cb_device_guard(int device) : dev(device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~cb_device_guard() {
        if (ok && prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
    cb_device_guard(const cb_device_guard&) = delete;
    cb_device_guard& operator=(const cb_device_guard&) = delete;
};

namespace cb {

// ------------------------------------------------------------------ tile-native activation layout
// Activations live in HBM the way the convolution's A operand wants them in shared memory, so a
// K-chunk of 64 channels of a board pair is ONE contiguous 25,600-byte bulk copy and the nine 3x3
// taps are nine descriptor offsets into it (no im2col, no TMA tensor map):
//   cell(P, j, R, b, Cc) = (((P*CJ + j)*10 + R)*2 + b)*10 + Cc      16 bytes = 8 bf16 channels 8j..8j+7
// P = board pair, j = 8-channel chunk (CJ = C/8), R = row+1, Cc = col+1 (zero border of one cell so
// "same" padding is free), b = board within the pair.  Interleaving the two boards row by row makes
// the stride between 8-pixel groups uniform (160 B) for all 16 groups of a 128-row MMA tile.
constexpr int ACT_CELLS = 200;          // 10 rows * 2 boards * 10 cols
constexpr int ACT_CHUNK_BYTES = 3200;   // ACT_CELLS * 16
constexpr int ACT_K64_BYTES = 25600;    // 8 chunks = 64 channels
__host__ __device__ inline int act_cell(int row, int board_in_pair, int col) {
    return ((row + 1) * 2 + board_in_pair) * 10 + (col + 1);
}
__host__ __device__ inline size_t act_offset_bytes(int pair, int cj_total, int j, int cell) {
    return ((size_t)(pair * cj_total + j) * ACT_CELLS + cell) * 16;
}
inline size_t act_bytes(int boards, int channels) { return (size_t)((boards + 1) / 2) * (channels / 8) * ACT_CHUNK_BYTES; }

#if defined(__CUDACC__)
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of a fully converged warp (always the same one): the warp runs the role's control flow together
// so that addresses and descriptors stay in uniform registers, and only the async instruction is predicated
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// the same for waits that are expected to last (a producer ahead of its consumer, epilogue warps waiting for the next accumulator, the
// flag publisher): the hardware may park the thread for up to `hint_ns` instead of returning at once, and wakes it when the phase
// completes.  A bare try_wait loop re-issues every few cycles (ncu: 35 M of a launch's 300 M warp instructions were such spins) and
// the tower runs against the power cap.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 20000) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy (UBLKCP), completion signalled on an mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster of 2 run one M=256 MMA together ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope), as CUTLASS does for the
// accumulator hand-back: a cluster-scope release would first drain this thread's outstanding global stores.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// bulk copy delivered to the same shared-memory offset (and signalling the barrier at the same offset) in every
// CTA of the cluster named in cta_mask
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// completion of this CTA's MMAs arrives on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// TMA tile load (2-D tensor map, rows of 128 B) into this CTA's shared memory; with .cta_group::2 the
// transaction bytes may complete on a barrier of the PEER CTA, given by its shared::cluster address
__device__ __forceinline__ void tma_load_2d_pair(void* dst_smem, const void* tensor_map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tensor_map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// TMA tile load (2-D tensor map) into this CTA's shared memory, completing on this CTA's barrier
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tensor_map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tensor_map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// the same tile delivered to the same shared-memory offset (and barrier offset) of every CTA of the cluster in cta_mask
__device__ __forceinline__ void tma_load_2d_multicast(void* dst_smem, const void* tensor_map, int c0, int c1, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst_smem)),
        "l"(tensor_map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tensor_map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tensor_map) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows from each CTA] * B[N/2 rows from each CTA]^T; issued by the leader CTA only
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same on e4m3 x e4m3 operands (K = 32 per MMA: the same 32 bytes per row as a bf16 K = 16 step).  Only the throughput probe of
// profiles/r2_precision_variants.md issues it (CB_TOWER_FP8_PROBE): e4m3 misses the 2e-2 tolerance by 4-7x, no datapath uses it.
__device__ __forceinline__ void umma2_e4m3(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all MMAs issued so far arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, one CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, no-swizzle ("interleave") shared-memory matrix descriptor: 8-row x 16-byte core matrices,
// LBO = byte stride between the two core matrices along K, SBO = byte stride between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both operands K-major
__device__ __forceinline__ uint32_t idesc_bf16_f32(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f8f6f4 instruction descriptor: e4m3 x e4m3 -> f32 (operand formats 0), both operands K-major
__device__ __forceinline__ uint32_t idesc_e4m3_f32(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
// the same with both operands MN-major (bits 15 / 16): shared memory holds 8 consecutive M (or N) indices in 16 bytes and
// steps along K from one 16-byte cell to the next.  In the no-swizzle descriptor LBO is then the stride between the
// 8-cell groups along K and SBO the stride between 8-index groups along M / N.
__device__ __forceinline__ uint32_t idesc_bf16_f32_mn(int m, int n) { return idesc_bf16_f32(m, n) | (1u << 15) | (1u << 16); }
}  // namespace ptx
#endif

}  // namespace cb
