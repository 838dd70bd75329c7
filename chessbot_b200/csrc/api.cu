// Device side of the C ABI (include/chessbot_b200.h): context, network weights and the batched
// search driver.  Kernels live in encode.cu / conv.cu / heads.cu / mcts.cu.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/chessbot_b200.h"
#include "kernels.cuh"

namespace cb {
__global__ void read_roots_kernel(SearchParams sp, uint16_t* moves, int32_t* n, float* w, float* q, double* p, int32_t* n_children,
                                  int32_t* root_n, int32_t* tree_nodes) {
    const int tree = blockIdx.x;
    if (tree >= sp.n_trees) return;
    const TreeState& ts = sp.trees[tree];
    const Node& r = sp.nodes[ts.root_node];
    if (threadIdx.x == 0) { n_children[tree] = r.n_children; root_n[tree] = r.N; tree_nodes[tree] = ts.n_nodes; }
    for (int i = threadIdx.x; i < r.n_children; i += blockDim.x) {
        const Node& c = sp.nodes[r.first_child + i];
        const size_t o = (size_t)tree * MAX_MOVES + i;
        moves[o] = c.move; n[o] = c.N; w[o] = c.W; q[o] = c.Q;
        p[o] = ts.noisy ? sp.root_p64[o] : (double)c.P;
    }
}
// compact variant: the children of tree t go to [child_off[t], child_off[t] + n_children) of every section (the caller knows the
// number of legal moves of every root); sections: P f64 | N i32 | W f32 | Q f32 | n_children, root_N, tree_nodes, state i32 [n] | moves u16
// (state = TreeState::status | error << 8)
__global__ void read_roots_compact_kernel(SearchParams sp, const int32_t* child_off, uint8_t* out, int total) {
    const int tree = blockIdx.x;
    if (tree >= sp.n_trees) return;
    const TreeState& ts = sp.trees[tree];
    const Node& r = sp.nodes[ts.root_node];
    double* P = reinterpret_cast<double*>(out);
    int32_t* N = reinterpret_cast<int32_t*>(P + total);
    float* W = reinterpret_cast<float*>(N + total);
    float* Q = W + total;
    int32_t* hdr = reinterpret_cast<int32_t*>(Q + total);
    uint16_t* mv = reinterpret_cast<uint16_t*>(hdr + 4 * sp.n_trees);
    const int off = child_off[tree], cap = child_off[tree + 1] - off;
    if (threadIdx.x == 0) {
        hdr[tree] = r.n_children; hdr[sp.n_trees + tree] = r.N; hdr[2 * sp.n_trees + tree] = ts.n_nodes;
        hdr[3 * sp.n_trees + tree] = ts.status | (ts.error ? 256 : 0);
    }
    const int k = min((int)r.n_children, cap);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const Node& c = sp.nodes[r.first_child + i];
        mv[off + i] = c.move; N[off + i] = c.N; W[off + i] = c.W; Q[off + i] = c.Q;
        P[off + i] = ts.noisy ? sp.root_p64[(size_t)tree * MAX_MOVES + i] : (double)c.P;
    }
}
__global__ void waiting_flags_kernel(const TreeState* trees, int n, uint8_t* flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = trees[i].status == 1 || trees[i].status == 2;
}
__global__ void count_unfinished_kernel(const TreeState* trees, int n, int32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && trees[i].status != 3 && trees[i].status != 0) atomicAdd(out, 1);
}
}  // namespace cb

using namespace cb;

// ------------------------------------------------------------------------------------ device memory
typedef void* (*cb_alloc_fn)(size_t bytes, int device, void* user);
typedef void (*cb_free_fn)(void* ptr, void* user);
static cb_alloc_fn g_alloc = nullptr;
static cb_free_fn g_free = nullptr;
static void* g_alloc_user = nullptr;
static bool g_detached = false;   // the host allocator is gone (interpreter shutting down): whatever is still held is left to process exit

int cb_device_alloc(void** ptr, size_t bytes) {
    *ptr = nullptr;
    if (!bytes) return 0;
    if (g_alloc) {
        int dev = 0;
        cudaGetDevice(&dev);
        *ptr = g_alloc(bytes, dev, g_alloc_user);
        if (!*ptr) { cb_set_error("host allocator returned no memory for " + std::to_string(bytes) + " bytes"); return -1; }
        return 0;
    }
    CB_CUDA_OK(cudaMalloc(ptr, bytes));
    return 0;
}
void cb_device_free(void* ptr) {
    if (!ptr || g_detached) return;
    if (g_free) g_free(ptr, g_alloc_user);
    else cudaFree(ptr);
}

namespace {
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count, bool zero = true) {
        free_();
        n = count;
        if (!count) return 0;
        if (cb_device_alloc((void**)&p, count * sizeof(T))) return -1;
        if (zero) CB_CUDA_OK(cudaMemset(p, 0, count * sizeof(T)));
        return 0;
    }
    void free_() { if (p) cb_device_free(p); p = nullptr; n = 0; }
    ~DevBuf() { free_(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {  // `net = Net()` on a reload frees the previous network's buffers
        if (this != &o) { free_(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
};

// pinned host staging owned by the context: asynchronous copies need no host synchronisation and no pageable bounce buffer
struct PinBuf {
    uint8_t* p = nullptr;
    size_t n = 0;
    int reserve(size_t bytes) {
        if (bytes <= n) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        bytes = bytes * 5 / 4 + 4096;
        CB_CUDA_OK(cudaMallocHost((void**)&p, bytes));
        n = bytes;
        return 0;
    }
    ~PinBuf() { if (p) cudaFreeHost(p); }
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
};

struct ConvLayer {
    DevBuf<__nv_bfloat16> w;
    DevBuf<float> scale, shift;
    int kc = 0, kc_in = 0, cj_in = 0;
};

struct Net {
    int filters = 0, cpad = 0, blocks = 0, max_batch = 0;
    float slope = 0.3f, eps = 1e-3f;
    std::map<std::string, std::vector<float>> host;  // staged Keras-layout tensors
    std::vector<ConvLayer> layers;                   // stem + 2 per block
    DevBuf<float> head_w, head_bn, fc1, fc2, wp, wpt;
    DevBuf<__nv_bfloat16> act[3], act_lo;            // act_lo: low half of the residual stream (stream_hi_lo)
    bool stream_hi_lo = true;
    int hi_lo_from_block = 0;                        // residual blocks before this one keep a bf16 stream (0: hi + lo everywhere)
    size_t l2_persist_bytes = 0;                     // persisting-L2 set-aside last requested for the residual stream's window
    DevBuf<float> head_raw, pfeat;
    DevBuf<LayerDesc> layer_desc, dbg_desc;          // tower kernel: per-layer descriptors (dbg_desc: cb_debug_conv_layer)
    DevBuf<CUtensorMap> wmaps;                       // tensor maps of the packed weights, one per layer
    DevBuf<int> flags;                               // per board pair progress flags of the tower kernel
    CUtensorMap act_maps[3];                         // tensor maps of act[0..2]
    CUtensorMap planes_map;                          // tensor map of the last planes buffer seen
    const void* planes_ptr = nullptr;
    int planes_pairs = 0;
    int epoch = 0;
    bool ready = false;
};

struct Search {
    int max_trees = 0, n_trees = 0, max_sims = 0, eval_mode = 0, leaves = 1;
    DevBuf<SlotState> slots;
    uint32_t seed = 0;
    DevBuf<Node> nodes;
    DevBuf<Pos> pos;
    DevBuf<TreeState> trees;
    DevBuf<int32_t> node_count, pos_count, frame_idx, counters, scratch_i32, path, active_list, active_count, slot_board;
    DevBuf<uint16_t> pend_moves;
    DevBuf<double> root_p64, noise, cpuct_tab, sqrt_tab;
    DevBuf<float> value, expvals;
    DevBuf<uint32_t> synth_hash;
    DevBuf<__nv_bfloat16> planes_act;
    DevBuf<int16_t> planes_i16;
    DevBuf<uint8_t> flags;
    // read-back staging
    DevBuf<uint16_t> r_moves; DevBuf<int32_t> r_n, r_nc, r_rootn, r_nodes; DevBuf<float> r_w, r_q; DevBuf<double> r_p;
    DevBuf<uint8_t> r_pack; DevBuf<int32_t> r_off;   // compact read-back: one packed device buffer, one D2H
    PinBuf h_begin, h_read;                          // pinned staging of cb_search_begin / cb_search_read_roots_compact_sync
    cudaEvent_t begin_copied = nullptr;              // the previous begin's H2D copies have left h_begin
    cudaEvent_t read_done = nullptr;                 // the packed read-back queued by cb_search_read_roots_compact_async has landed in h_read
    size_t read_bytes = 0, read_cnt_off = 0;
    SearchParams sp{};
};
}  // namespace

struct cb_ctx {
    int device = 0, sms = 0;
    bool profile = false;                       // CUDA events around every kernel launch (bench.py roofline legs)
    std::vector<cudaEvent_t> prof_events;       // pairs
    std::vector<int> prof_stage;                // CB_PROF_* of each pair
    size_t prof_used = 0;
    DevBuf<float> exp_table;
    Net net;
    Search search;
    // cb_leaf_eval_host_sync: device scratch (grow-only) and pinned staging
    struct LeafEval {
        int cap = 0;
        DevBuf<uint8_t> pos, pack;
        DevBuf<int32_t> cnt;
        DevBuf<__nv_bfloat16> planes;
        DevBuf<float> value, priors;
        DevBuf<uint16_t> moves;
        DevBuf<int16_t> actions;
        PinBuf h_in, h_out;
    } leaf;
};

// brackets one kernel launch with CUDA events on its stream when profiling is on
static int prof_begin(cb_ctx* ctx, int stage, cudaStream_t st) {
    if (!ctx->profile) return 0;
    while (ctx->prof_events.size() < ctx->prof_used + 2) {
        cudaEvent_t e;
        CB_CUDA_OK(cudaEventCreate(&e));
        ctx->prof_events.push_back(e);
    }
    if (ctx->prof_stage.size() < ctx->prof_used / 2 + 1) ctx->prof_stage.resize(ctx->prof_used / 2 + 1);
    ctx->prof_stage[ctx->prof_used / 2] = stage;
    CB_CUDA_OK(cudaEventRecord(ctx->prof_events[ctx->prof_used], st));
    return 0;
}
static int prof_end(cb_ctx* ctx, cudaStream_t st) {
    if (!ctx->profile) return 0;
    CB_CUDA_OK(cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], st));
    ctx->prof_used += 2;
    return 0;
}

// every entry point runs on its context's device, whatever the caller's current device is (several contexts per process)
#define CB_ENTER(ctx)                                                                                                  \
    if (!(ctx)) { cb_set_error("null context"); return -1; }                                                           \
    cb_device_guard cb_guard_((ctx)->device);                                                                          \
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(context device) failed"); return -1; }

static int upload(void* dst, const void* src, size_t bytes) {
    CB_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" {

int cb_set_device_allocator(void* (*alloc)(size_t bytes, int device, void* user), void (*release)(void* ptr, void* user), void* user) {
    if ((alloc == nullptr) != (release == nullptr)) { cb_set_error("cb_set_device_allocator: give both functions or neither"); return -1; }
    if (!alloc && g_alloc) g_detached = true;   // blocks handed out by the host allocator must not reach cudaFree
    g_alloc = alloc;
    g_free = release;
    g_alloc_user = user;
    return 0;
}
// bytes cb_search_create(max_trees, max_nodes, max_positions, ...) + cb_search_set_leaves(leaves) take from the allocator
size_t cb_search_workspace_bytes(int max_trees, int max_nodes, int max_positions, int max_sims, int leaves) {
    const size_t mt = max_trees, mm = mt * MAX_MOVES, ns = mt * (size_t)leaves;
    size_t b = ((size_t)max_nodes + mt * NODE_CHUNK) * sizeof(Node) + ((size_t)max_positions + mt * POS_CHUNK) * sizeof(Pos) + mt * sizeof(TreeState);
    b += mm * (8 + 8) + mt;                                                       // root priors f64, noise, flags
    b += mm * (2 + 4 + 4 + 4 + 8) + mt * 12;                                      // dense read-back staging
    b += cb_search_roots_compact_bytes(max_trees, (int)mm) + (mt + 1) * 4;        // packed read-back staging
    b += (size_t)(max_sims + 8) * 16;                                             // ucb tables
    b += ns * (4 + 8 * 4 + MCTS_MAX_PATH * 4 + MAX_MOVES * 2 + 4 + 4 + sizeof(SlotState)) + act_bytes((int)ns, 128) + ns * 7616 * 2;
    return b;
}

cb_ctx* cb_ctx_create(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cb_set_error("no CUDA device: chessbot_b200 has no CPU fallback");
        return nullptr;
    }
    cb_device_guard guard(device);
    if (!guard.ok) { cb_set_error("cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        cb_set_error(std::string("chessbot_b200 needs an sm_100 GPU, found sm_") + std::to_string(prop.major * 10 + prop.minor));
        return nullptr;
    }
    cb_ctx* ctx = new cb_ctx;
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    std::vector<float> tab(65536);
    for (uint32_t i = 0; i < 65536; ++i) {
        uint32_t bits = i << 16;
        float x;
        memcpy(&x, &bits, 4);
        tab[i] = (float)exp((double)x);
    }
    if (ctx->exp_table.alloc(65536) != 0 || upload(ctx->exp_table.p, tab.data(), 65536 * 4) != 0) { delete ctx; return nullptr; }
    return ctx;
}
void cb_ctx_destroy(cb_ctx* ctx) {
    if (!ctx) return;
    cb_device_guard guard(ctx->device);
    delete ctx;
}
int cb_ctx_sm_count(const cb_ctx* ctx) { return ctx->sms; }
int cb_ctx_device(const cb_ctx* ctx) { return ctx->device; }
int cb_set_exp_table(cb_ctx* ctx, const float* table_host) {
    CB_ENTER(ctx);
    return upload(ctx->exp_table.p, table_host, 65536 * 4);
}

size_t cb_planes_act_bytes(int n) { return act_bytes(n, 128); }
int cb_encode(cb_ctx* ctx, const void* pos_dev, const int32_t* frame_idx_dev, int n, void* planes_act_dev, int16_t* planes_i16_dev, void* stream) {
    CB_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    if (prof_begin(ctx, CB_PROF_ENCODE, st) ||
        launch_encode((const Pos*)pos_dev, frame_idx_dev, n, (__nv_bfloat16*)planes_act_dev, planes_i16_dev, st) || prof_end(ctx, st))
        return -1;
    return 0;
}

// ------------------------------------------------------------------------------------ network
int cb_net_create(cb_ctx* ctx, int filters, int blocks, int max_batch, float leaky_slope, float bn_eps) {
    if (filters < 1 || filters > 256 || blocks < 0 || max_batch < 1) { cb_set_error("cb_net_create: bad shape (filters 1..256)"); return -1; }
    Net& net = ctx->net;
    net = Net();
    net.filters = filters;
    net.cpad = (filters + 63) / 64 * 64;
    net.blocks = blocks;
    net.max_batch = max_batch;
    net.slope = leaky_slope;
    net.eps = bn_eps;
    return 0;
}
int cb_net_set_stream_precision(cb_ctx* ctx, int hi_lo) {
    // hi_lo: 0 = bf16 stream, 1 = hi + lo from the stem on, k >= 2 = hi + lo from the output of residual block k - 1 on
    ctx->net.stream_hi_lo = hi_lo != 0;
    ctx->net.hi_lo_from_block = hi_lo >= 2 ? hi_lo - 1 : 0;
    ctx->net.ready = false;
    return 0;
}
int cb_net_set_tensor(cb_ctx* ctx, const char* name, const float* host, size_t count) {
    ctx->net.host[name] = std::vector<float>(host, host + count);
    ctx->net.ready = false;
    return 0;
}

static int pack_conv(Net& net, ConvLayer& L, const std::vector<float>& w, int cin, int cin_pad, const std::vector<float>& bn, bool stem) {
    const int C = net.filters, N = net.cpad;
    if (w.size() != (size_t)9 * cin * C || bn.size() != (size_t)4 * C) { cb_set_error("conv tensor has wrong size"); return -1; }
    // The stem sees raw integer planes (move count, halfmove clock up to the hundreds): its weights are
    // split into bf16 hi + lo halves (two K passes over the same planes) so they act at ~16-bit precision.
    const int halves = stem ? 2 : 1;
    L.kc_in = cin_pad / 64;
    L.kc = L.kc_in * halves;
    L.cj_in = cin_pad / 8;
    std::vector<__nv_bfloat16> packed((size_t)L.kc * 9 * N * 64, __float2bfloat16(0.f));
    for (int tap = 0; tap < 9; ++tap)
        for (int ci = 0; ci < cin_pad; ++ci) {
            int src_ci = ci;
            if (stem) {  // channels 119 / 120 carry the bf16 remainders of move count / halfmove clock (encode.cu)
                if (ci == 119) src_ci = 113;
                else if (ci == 120) src_ci = 118;
                else if (ci > 120) continue;
            } else if (ci >= cin) continue;
            const int k = ci & 63;
            for (int co = 0; co < C; ++co) {
                const float wf = w[((size_t)tap * cin + src_ci) * C + co];
                const __nv_bfloat16 hi = __float2bfloat16(wf);
                for (int h = 0; h < halves; ++h) {
                    const int kc = (ci >> 6) + h * L.kc_in;
                    packed[conv_weight_index(kc, tap, k, co, N)] =
                        h == 0 ? hi : __float2bfloat16(wf - __bfloat162float(hi));
                }
            }
        }
    std::vector<float> scale(N, 0.f), shift(N, 0.f);
    for (int c = 0; c < C; ++c) {
        const double g = bn[c], b = bn[C + c], m = bn[2 * C + c], v = bn[3 * C + c];
        const double s = g / sqrt(v + (double)net.eps);
        scale[c] = (float)s;
        shift[c] = (float)(b - m * s);
    }
    if (L.w.alloc(packed.size(), false) || L.scale.alloc(N, false) || L.shift.alloc(N, false)) return -1;
    if (upload(L.w.p, packed.data(), packed.size() * 2) || upload(L.scale.p, scale.data(), N * 4) || upload(L.shift.p, shift.data(), N * 4)) return -1;
    return 0;
}

int cb_net_finalize(cb_ctx* ctx) {
    CB_ENTER(ctx);
    Net& net = ctx->net;
    const int C = net.filters, N = net.cpad;
    auto need = [&](const std::string& k) -> const std::vector<float>* {
        auto it = net.host.find(k);
        if (it == net.host.end()) { cb_set_error("missing tensor: " + k); return nullptr; }
        return &it->second;
    };
    net.layers.clear();
    net.layers.resize(1 + 2 * net.blocks);
    const std::vector<float>*w = nullptr, *bn = nullptr;
    if (!(w = need("stem.w")) || !(bn = need("stem.bn"))) return -1;
    if (pack_conv(net, net.layers[0], *w, 119, 128, *bn, true)) return -1;
    for (int i = 0; i < net.blocks; ++i)
        for (int h = 0; h < 2; ++h) {
            const std::string base = "res" + std::to_string(i);
            if (!(w = need(base + (h ? ".w2" : ".w1"))) || !(bn = need(base + (h ? ".bn2" : ".bn1")))) return -1;
            if (pack_conv(net, net.layers[1 + 2 * i + h], *w, C, N, *bn, false)) return -1;
        }
    const std::vector<float>*vc = nullptr, *vbn = nullptr, *fc1 = nullptr, *fc2 = nullptr, *pc = nullptr, *pbn = nullptr, *pfc = nullptr;
    if (!(vc = need("value.conv")) || !(vbn = need("value.bn")) || !(fc1 = need("value.fc1")) || !(fc2 = need("value.fc2")) ||
        !(pc = need("policy.conv")) || !(pbn = need("policy.bn")) || !(pfc = need("policy.fc")))
        return -1;
    if (vc->size() != (size_t)C || pc->size() != (size_t)2 * C || fc1->size() != (size_t)64 * C || fc2->size() != (size_t)C ||
        pfc->size() != (size_t)128 * 4672 || vbn->size() != 4 || pbn->size() != 8) { cb_set_error("head tensor has wrong size"); return -1; }
    std::vector<float> head_w(3 * N, 0.f), head_bn(6);
    for (int c = 0; c < C; ++c) {
        head_w[c] = (*vc)[c];
        head_w[N + c] = (*pc)[c * 2];
        head_w[2 * N + c] = (*pc)[c * 2 + 1];
    }
    {
        const double s = (*vbn)[0] / sqrt((*vbn)[3] + (double)net.eps);
        head_bn[0] = (float)s;
        head_bn[1] = (float)((*vbn)[1] - (*vbn)[2] * s);
        for (int c = 0; c < 2; ++c) {
            const double sp = (*pbn)[c] / sqrt((*pbn)[6 + c] + (double)net.eps);
            head_bn[2 + 2 * c] = (float)sp;
            head_bn[3 + 2 * c] = (float)((*pbn)[2 + c] - (*pbn)[4 + c] * sp);
        }
    }
    std::vector<float> wpt((size_t)4672 * 128);
    for (int k = 0; k < 128; ++k)
        for (int a = 0; a < 4672; ++a) wpt[(size_t)a * 128 + k] = (*pfc)[(size_t)k * 4672 + a];
    if (net.head_w.alloc(3 * N, false) || net.head_bn.alloc(6, false) || net.fc1.alloc(64 * C, false) || net.fc2.alloc(C, false) ||
        net.wp.alloc((size_t)128 * 4672, false) || net.wpt.alloc((size_t)128 * 4672, false))
        return -1;
    if (upload(net.head_w.p, head_w.data(), head_w.size() * 4) || upload(net.head_bn.p, head_bn.data(), 24) ||
        upload(net.fc1.p, fc1->data(), fc1->size() * 4) || upload(net.fc2.p, fc2->data(), fc2->size() * 4) ||
        upload(net.wp.p, pfc->data(), pfc->size() * 4) || upload(net.wpt.p, wpt.data(), wpt.size() * 4))
        return -1;
    const size_t act_elems = act_bytes(net.max_batch, N) / 2;
    for (int i = 0; i < 2; ++i)
        if (net.act[i].alloc(act_elems)) return -1;  // zeroed: border cells stay zero forever
    net.act[2].free_();
    if (net.stream_hi_lo ? net.act_lo.alloc(act_elems) : (net.act_lo.free_(), 0)) return -1;
    const size_t mb = (size_t)(net.max_batch + 3) / 4 * 4;
    if (net.head_raw.alloc(mb * 192) || net.pfeat.alloc(mb * 128)) return -1;
    // tower kernel: weight tensor maps, layer descriptors (buffer 0 = input planes, 1..3 = act[0..2]), progress flags
    const int nl = (int)net.layers.size();
    std::vector<CUtensorMap> wm(nl);
    std::vector<LayerDesc> ld(nl);
    for (int l = 0; l < nl; ++l) {
        ConvLayer& L = net.layers[l];
        if (conv_make_row_map(&wm[l], L.w.p, (uint64_t)L.kc * 9 * N, conv_taps_per_box(N) * N / 2)) return -1;
        LayerDesc& d = ld[l];
        d.scale = L.scale.p; d.shift = L.shift.p; d.kc = L.kc; d.kc_in = L.kc_in; d.cj_in = L.cj_in; d.w_map = l; d.pad_ = 0;
        // Two activation buffers: 1 = the residual stream, 2 = the block's inner activation.  The second convolution of a block writes the new
        // stream value over the old one IN PLACE: the epilogue thread that reads a cell of the skip tile is the thread that stores that cell of
        // the output, and nothing else reads the old stream any more (its first convolution is done).  One tensor less in the working set
        // (3 x 52 MB instead of 4 at 1024 boards of 20x256 incl. the lo half), and the store hits a line the skip load has just brought in.
        if (l == 0) { d.in_buf = 0; d.out_buf = 1; d.skip_buf = -1; }
        else if (l & 1) { d.in_buf = 1; d.out_buf = 2; d.skip_buf = -1; }   // first conv of a block
        else { d.in_buf = 2; d.out_buf = 1; d.skip_buf = 1; }               // second conv + skip
    }
    if (net.wmaps.alloc(nl, false) || net.layer_desc.alloc(nl, false) || net.dbg_desc.alloc(1, false) ||
        net.flags.alloc(2 * ((size_t)net.max_batch / 4 + 2)))
        return -1;
    if (upload(net.wmaps.p, wm.data(), nl * sizeof(CUtensorMap)) || upload(net.layer_desc.p, ld.data(), nl * sizeof(LayerDesc))) return -1;
    for (int i = 0; i < 2; ++i)
        if (conv_make_row_map(&net.act_maps[i], net.act[i].p, (uint64_t)act_bytes(net.max_batch, N) / 128, ACT_K64_BYTES / 128)) return -1;
    net.planes_ptr = nullptr;
    net.epoch = 0;
    net.host.clear();
    net.ready = true;
    return 0;
}

// progress flags hold epoch * 256 + layers finished; start a new epoch per launch, clear before it can overflow
static int next_epoch(Net& net, cudaStream_t st) {
    if (++net.epoch >= (1 << 22)) {
        CB_CUDA_OK(cudaMemsetAsync(net.flags.p, 0, net.flags.n * sizeof(int), st));
        net.epoch = 1;
    }
    return 0;
}

// tower + head features; leaves pfeat/value for the callers below
// slot_list / n_dev (search loop): the batch holds *n_dev <= n boards, board b belongs to evaluation slot slot_list[b]
static int net_tower(cb_ctx* ctx, const void* planes_act_dev, int n, float* value_dev, cudaStream_t st, const int32_t* slot_list = nullptr,
                     const int* n_dev = nullptr, bool run_heads = true) {
    Net& net = ctx->net;
    if (!net.ready) { cb_set_error("network not finalized"); return -1; }
    if (n > net.max_batch) { cb_set_error("batch larger than max_batch"); return -1; }
    const int pairs = (n + 1) / 2;
    if (net.planes_ptr != planes_act_dev || net.planes_pairs != pairs) {
        if (conv_make_row_map(&net.planes_map, planes_act_dev, (uint64_t)act_bytes(n, 128) / 128, ACT_K64_BYTES / 128)) return -1;
        net.planes_ptr = planes_act_dev;
        net.planes_pairs = pairs;
    }
    if (next_epoch(net, st)) return -1;
    TowerParams tp{};
    tp.tm_act[0] = net.planes_map;
    tp.bufs[0] = (uint8_t*)const_cast<void*>(planes_act_dev);
    for (int i = 0; i < 2; ++i) { tp.tm_act[1 + i] = net.act_maps[i]; tp.bufs[1 + i] = (uint8_t*)net.act[i].p; }
    tp.maps = net.wmaps.p;
    tp.layers = net.layer_desc.p;
    tp.head_w = net.head_w.p;
    tp.head_out = net.head_raw.p;
    tp.lo_buf = (uint8_t*)net.act_lo.p;
    {
        // measurement only (profiles/r2_precision_variants.md): bit 0 = e4m3 throughput probe; CB_TOWER_DIAG bits 1 / 2 = the hi + lo stream
        // without its lo stores / without its lo loads (numerically meaningless: where does its cost come from?)
        static const int probe = (getenv("CB_TOWER_FP8_PROBE") ? 1 : 0) | (getenv("CB_TOWER_DIAG") ? 2 * atoi(getenv("CB_TOWER_DIAG")) : 0);
        tp.fp8_probe = probe;
    }
    tp.lo_from_layer = net.hi_lo_from_block > 0 ? 2 * net.hi_lo_from_block : 0;  // block b's output is layer 2 b + 2; its input stream is layer 2 b
    tp.n_boards_dev = n_dev;
    tp.full_sectors = (net.stream_hi_lo ? 3 : 2) * act_bytes(n, net.cpad) > ((size_t)96 << 20);  // stream + inner activation (+ lo) do not fit the 126 MB L2
    tp.flags = net.flags.p;
    tp.epoch = net.epoch;
    tp.nl = (int)net.layers.size();
    tp.pairs = pairs;
    tp.n = net.cpad;
    tp.slope = net.slope;
    if (tp.full_sectors) {
        // When the activations overflow L2 every tensor makes a round trip through DRAM between the layer that writes it and the layers that
        // read it.  The residual stream is the most re-used one (written by a block's second convolution in place, read by the next block's
        // first convolution and again as the skip tile): a persisting access-policy window on the launching stream keeps it L2-resident
        // (set-aside = its size + 4 MB, at most 64 MB of the 126 MB).  Measured at 1024 boards of 20x256 (profiles/README.md): tower -1.5 %,
        // +1.3 % evals/s at the power cap; the window on the lo half instead: +0.5 %.  CB_TOWER_L2_PERSIST=0 turns it off, =<MB> overrides the size.
        static const int persist_env = getenv("CB_TOWER_L2_PERSIST") ? atoi(getenv("CB_TOWER_L2_PERSIST")) : -1;
        const size_t win_bytes = act_bytes(n, net.cpad);
        const size_t want = persist_env >= 0 ? (size_t)persist_env << 20 : ((win_bytes + ((size_t)4 << 20)) < ((size_t)64 << 20) ? win_bytes + ((size_t)4 << 20) : (size_t)64 << 20);
        if (want > 0) {
            if (net.l2_persist_bytes != want) {
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) cudaGetLastError();  // not fatal: the window then has no effect
                net.l2_persist_bytes = want;
            }
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = net.act[0].p;
            av.accessPolicyWindow.num_bytes = win_bytes;
            av.accessPolicyWindow.hitRatio = win_bytes <= want ? 1.0f : (float)((double)want / (double)win_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
        }
    }
    if (prof_begin(ctx, CB_PROF_CONV, st) || launch_tower(tp, ctx->sms, st) || prof_end(ctx, st)) return -1;
    if (!run_heads) return 0;  // the search finishes the heads inside mcts_step (SearchParams::head_raw)
    HeadParams hp{net.head_raw.p, net.head_bn.p, net.fc1.p, net.fc2.p, net.filters, net.slope, value_dev, net.pfeat.p, slot_list, n_dev};
    if (prof_begin(ctx, CB_PROF_HEADS, st) || launch_head_features(hp, n, st) || prof_end(ctx, st)) return -1;
    return 0;
}

int cb_net_forward(cb_ctx* ctx, const void* planes_act_dev, int n, float* value_dev, void* logits_bf16_dev, float* logits_f32_dev, void* stream) {
    CB_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    if (net_tower(ctx, planes_act_dev, n, value_dev, st)) return -1;
    if (logits_bf16_dev || logits_f32_dev) {
        if (prof_begin(ctx, CB_PROF_POLICY_DENSE, st) ||
            launch_policy_dense(ctx->net.pfeat.p, ctx->net.wp.p, n, (__nv_bfloat16*)logits_bf16_dev, logits_f32_dev, st) || prof_end(ctx, st))
            return -1;
    }
    return 0;
}

int cb_synth_forward(cb_ctx* ctx, const int16_t* planes_i16_dev, int n, uint32_t seed, float* value_dev, void* logits_bf16_dev, void* stream) {
    CB_ENTER(ctx);
    (void)ctx;
    static thread_local DevBuf<uint32_t> hash;  // scratch of the stand-alone call (the search owns its own per-slot hashes)
    if (hash.n < (size_t)n && hash.alloc(n)) return -1;
    return launch_synth(planes_i16_dev, n, seed, hash.p, value_dev, (__nv_bfloat16*)logits_bf16_dev, (cudaStream_t)stream);
}

int cb_policy_softmax(cb_ctx* ctx, const void* pos_dev, const int32_t* cur_idx_dev, int n, const void* logits_bf16_dev,
                      uint16_t* out_moves_dev, int16_t* out_actions_dev, float* out_priors_dev, int32_t* out_n_dev, void* stream) {
    CB_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    if (prof_begin(ctx, CB_PROF_POLICY_SOFTMAX, st) ||
        launch_policy_softmax((const Pos*)pos_dev, cur_idx_dev, n, (const __nv_bfloat16*)logits_bf16_dev, ctx->exp_table.p,
                              out_moves_dev, out_actions_dev, out_priors_dev, out_n_dev, st) || prof_end(ctx, st))
        return -1;
    return 0;
}

// ---- BASELINE config 2: batched leaf evaluation outside a tree (encode + forward + legal-move mask / softmax) -------------
int cb_leaf_eval(cb_ctx* ctx, const void* pos_dev, const int32_t* frame_idx_dev, const int32_t* cur_idx_dev, int n, void* planes_act_dev,
                 float* value_dev, uint16_t* moves_dev, int16_t* actions_dev, float* priors_dev, int32_t* n_moves_dev, void* stream) {
    CB_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    if (prof_begin(ctx, CB_PROF_ENCODE, st) || launch_encode((const Pos*)pos_dev, frame_idx_dev, n, (__nv_bfloat16*)planes_act_dev, nullptr, st) ||
        prof_end(ctx, st))
        return -1;
    if (net_tower(ctx, planes_act_dev, n, value_dev, st)) return -1;
    if (prof_begin(ctx, CB_PROF_POLICY_SOFTMAX, st) ||
        launch_leaf_priors((const Pos*)pos_dev, cur_idx_dev, n, ctx->net.pfeat.p, ctx->net.wpt.p, ctx->exp_table.p, moves_dev, actions_dev, priors_dev,
                           n_moves_dev, st) ||
        prof_end(ctx, st))
        return -1;
    return 0;
}
size_t cb_leaf_eval_packed_bytes(int n, int total_moves) { return (size_t)n * 12 + 4 + (size_t)total_moves * 8; }
// The same from HOST game records, end to end: records -> pinned staging -> H2D -> kernels -> packed results -> ONE small + one sized
// device-to-host copy.  out_host receives value f32[n] | n_moves i32[n] | offsets i32[n + 1] | priors f32[T] | moves u16[T] | actions i16[T]
// (T = offsets[n] <= total legal moves); returns T, or -1.  out_cap_bytes must hold cb_leaf_eval_packed_bytes(n, T).
int cb_leaf_eval_host_sync(cb_ctx* ctx, const void* records_host, const int32_t* rec_len, int n, void* out_host, size_t out_cap_bytes, void* stream) {
    CB_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    auto& L = ctx->leaf;
    if (n < 1) { cb_set_error("cb_leaf_eval_host_sync: n < 1"); return -1; }
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += rec_len[i];
    const int cap_moves = n * MAX_MOVES;
    if (n > L.cap) {
        const size_t nn = n;
        if (L.cnt.alloc(nn, false) || L.planes.alloc(act_bytes(n, 128) / 2) ||
            L.value.alloc(nn, false) || L.priors.alloc(nn * MAX_MOVES, false) || L.moves.alloc(nn * MAX_MOVES, false) ||
            L.actions.alloc(nn * MAX_MOVES, false) || L.pack.alloc(cb_leaf_eval_packed_bytes(n, cap_moves), false))
            return -1;
        L.cap = n;
    }
    // staging: positions | frame_idx [n][8] | cur_idx [n], mirrored by ONE device block (one H2D)
    const size_t b_pos = total * sizeof(Pos), o_fi = (b_pos + 15) & ~(size_t)15, o_ci = o_fi + (size_t)n * 32, b_in = o_ci + (size_t)n * 4;
    if (L.pos.n < b_in && L.pos.alloc(b_in * 5 / 4 + 4096, false)) return -1;
    if (L.h_in.reserve(b_in)) return -1;
    memcpy(L.h_in.p, records_host, b_pos);
    const Pos* rec = reinterpret_cast<const Pos*>(L.h_in.p);
    int32_t* fi = reinterpret_cast<int32_t*>(L.h_in.p + o_fi);
    int32_t* ci = reinterpret_cast<int32_t*>(L.h_in.p + o_ci);
    size_t off = 0;
    for (int i = 0; i < n; ++i) {  // frames newest first: back to the start of the record or of the game (gameimage.py:170-181)
        int cur = (int)(off + rec_len[i] - 1);
        ci[i] = cur;
        for (int t = 0; t < 8; ++t) fi[i * 8 + t] = -1;
        for (int t = 0; t < 8; ++t) {
            fi[i * 8 + t] = cur;
            if (rec[cur].ply == 0 || cur == (int)off) break;
            --cur;
        }
        off += rec_len[i];
    }
    CB_CUDA_OK(cudaMemcpyAsync(L.pos.p, L.h_in.p, b_in, cudaMemcpyHostToDevice, st));
    const int32_t* d_fi = reinterpret_cast<const int32_t*>(L.pos.p + o_fi);
    const int32_t* d_ci = reinterpret_cast<const int32_t*>(L.pos.p + o_ci);
    if (cb_leaf_eval(ctx, L.pos.p, d_fi, d_ci, n, L.planes.p, L.value.p, L.moves.p, L.actions.p, L.priors.p, L.cnt.p, stream)) return -1;
    if (launch_pack_leaf_eval(n, L.value.p, L.cnt.p, L.moves.p, L.actions.p, L.priors.p, L.pack.p, st)) return -1;
    // one speculative read (header + 64 moves per position: the average is ~30), a second one only if a batch needs more
    const size_t b_hdr = (size_t)n * 12 + 4;
    const size_t b_all = cb_leaf_eval_packed_bytes(n, cap_moves), b_first = b_hdr + (size_t)n * 64 * 8 < b_all ? b_hdr + (size_t)n * 64 * 8 : b_all;
    if (L.h_out.reserve(b_all)) return -1;
    CB_CUDA_OK(cudaMemcpyAsync(L.h_out.p, L.pack.p, b_first, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));
    const int T = reinterpret_cast<const int32_t*>(L.h_out.p)[3 * n];
    if (T < 0 || T > cap_moves) { cb_set_error("cb_leaf_eval_host_sync: corrupt offsets"); return -1; }
    const size_t b_need = cb_leaf_eval_packed_bytes(n, T);
    if (b_need > out_cap_bytes) { cb_set_error("cb_leaf_eval_host_sync: output buffer too small"); return -1; }
    if (b_need > b_first) {
        CB_CUDA_OK(cudaMemcpyAsync(L.h_out.p + b_first, L.pack.p + b_first, b_need - b_first, cudaMemcpyDeviceToHost, st));
        CB_CUDA_OK(cudaStreamSynchronize(st));
    }
    memcpy(out_host, L.h_out.p, b_need);
    return T;
}

int cb_profile(cb_ctx* ctx, int enable) {
    ctx->profile = enable != 0;
    ctx->prof_used = 0;
    return 0;
}
int cb_profile_read_sync(cb_ctx* ctx, double* total_ms, int* launches) {
    CB_ENTER(ctx);
    CB_CUDA_OK(cudaDeviceSynchronize());
    for (int k = 0; k < CB_PROF_STAGES; ++k) { total_ms[k] = 0; launches[k] = 0; }
    for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
        float ms = 0;
        CB_CUDA_OK(cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]));
        total_ms[ctx->prof_stage[i / 2]] += ms;
        launches[ctx->prof_stage[i / 2]] += 1;
    }
    ctx->prof_used = 0;
    return 0;
}

int cb_debug_conv_layer(cb_ctx* ctx, int layer, const void* in_act_dev, const void* skip_act_dev, void* out_act_dev, int n_boards,
                        int use_reference, void* stream) {
    CB_ENTER(ctx);
    Net& net = ctx->net;
    cudaStream_t st = (cudaStream_t)stream;
    if (!net.ready || layer < 0 || layer >= (int)net.layers.size()) { cb_set_error("bad layer"); return -1; }
    if (n_boards > net.max_batch) { cb_set_error("batch larger than max_batch"); return -1; }
    ConvLayer& L = net.layers[layer];
    const int pairs = (n_boards + 1) / 2;
    if (use_reference) {
        ConvParams p{};
        p.in = (const __nv_bfloat16*)in_act_dev; p.out = (__nv_bfloat16*)out_act_dev; p.skip = (const __nv_bfloat16*)skip_act_dev;
        p.wgt = L.w.p; p.scale = L.scale.p; p.shift = L.shift.p;
        p.pairs = pairs; p.kc = L.kc; p.kc_in = L.kc_in; p.cj_in = L.cj_in; p.n = net.cpad; p.slope = net.slope;
        return launch_conv3x3_reference(p, st);
    }
    // the tower kernel with a one-layer plan: buffer 0 = in, 1 = out, 2 = skip
    LayerDesc d{};
    d.scale = L.scale.p; d.shift = L.shift.p; d.kc = L.kc; d.kc_in = L.kc_in; d.cj_in = L.cj_in; d.w_map = layer;
    d.in_buf = 0; d.out_buf = 1; d.skip_buf = skip_act_dev ? 2 : -1;
    CB_CUDA_OK(cudaMemcpyAsync(net.dbg_desc.p, &d, sizeof(d), cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));  // `d` is on this stack frame
    if (next_epoch(net, st)) return -1;
    TowerParams tp{};
    if (conv_make_row_map(&tp.tm_act[0], in_act_dev, (uint64_t)pairs * L.cj_in * (ACT_CHUNK_BYTES / 128), ACT_K64_BYTES / 128)) return -1;
    tp.bufs[0] = (uint8_t*)const_cast<void*>(in_act_dev);
    tp.bufs[1] = (uint8_t*)out_act_dev;
    tp.bufs[2] = (uint8_t*)const_cast<void*>(skip_act_dev);
    tp.maps = net.wmaps.p;
    tp.layers = net.dbg_desc.p;
    tp.flags = net.flags.p;
    tp.epoch = net.epoch;
    tp.nl = 1;
    tp.pairs = pairs;
    tp.n = net.cpad;
    tp.slope = net.slope;
    return launch_tower(tp, ctx->sms, st);
}

// weight gradient of one 3x3 convolution on tile-native buffers: out_dev f32 [9][cin_pad][n] (padded channels included)
int cb_debug_wgrad(cb_ctx* ctx, const void* in_act_dev, const void* dy_act_dev, int n_boards, int cin_pad, int n, float* out_dev,
                   int use_reference, void* stream) {
    CB_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    const int pairs = (n_boards + 1) / 2, cj_in = cin_pad / 8;
    if (use_reference)
        return launch_wgrad_reference((const __nv_bfloat16*)in_act_dev, (const __nv_bfloat16*)dy_act_dev, pairs, cj_in, n, out_dev, st);
    WgradParams wp{};
    wp.pairs = pairs; wp.cj_in = cj_in; wp.n = n; wp.splits = wgrad_splits(ctx->sms, cj_in);
    const int dy_chunks = n / 16;
    if (conv_make_row_map(&wp.tm_a, in_act_dev, (uint64_t)pairs * cj_in * 25, 200) ||
        conv_make_row_map(&wp.tm_dy, dy_act_dev, (uint64_t)pairs * (n / 8) * 25, 25 * (dy_chunks < 8 ? dy_chunks : 8)))
        return -1;
    DevBuf<float> partial;
    if (partial.alloc((size_t)wp.splits * 9 * cin_pad * n, false)) return -1;
    wp.partial = partial.p;
    if (launch_wgrad(wp, st) || launch_wgrad_reduce(partial.p, wp.splits, cin_pad, n, cin_pad, n, 0, out_dev, st)) return -1;
    CB_CUDA_OK(cudaStreamSynchronize(st));  // the scratch buffer is freed on return
    return 0;
}

// ------------------------------------------------------------------------------------ search
int cb_search_create(cb_ctx* ctx, int max_trees, int max_nodes, int max_positions, int max_sims) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    s.max_trees = max_trees;
    s.max_sims = max_sims;
    // every tree allocates from the pools in chunks: the last chunk of each tree may stay partly unused
    max_nodes += max_trees * NODE_CHUNK;
    max_positions += max_trees * POS_CHUNK;
    const size_t mt = max_trees, mm = (size_t)max_trees * MAX_MOVES;
    if (s.nodes.alloc(max_nodes, false) || s.pos.alloc(max_positions, false) || s.trees.alloc(mt) || s.node_count.alloc(1) ||
        s.pos_count.alloc(1) || s.counters.alloc(4) || s.scratch_i32.alloc(4) ||
        s.root_p64.alloc(mm) || s.noise.alloc(mm) || s.flags.alloc(mt) ||
        s.r_moves.alloc(mm) || s.r_n.alloc(mm) || s.r_w.alloc(mm) || s.r_q.alloc(mm) || s.r_p.alloc(mm) || s.r_nc.alloc(mt) ||
        s.r_rootn.alloc(mt) || s.r_nodes.alloc(mt))
        return -1;
    // ucb constants from the host libm, like the reference computes them (mcts.py:136-140)
    const int tl = max_sims + 8;
    std::vector<double> ct(tl), sq(tl);
    for (int n = 0; n < tl; ++n) {
        ct[n] = log((double)(n + 19652 + 1) / 19652.0) + 1.25;
        sq[n] = pow((double)n, 0.5);
    }
    if (s.cpuct_tab.alloc(tl, false) || s.sqrt_tab.alloc(tl, false) || upload(s.cpuct_tab.p, ct.data(), tl * 8) ||
        upload(s.sqrt_tab.p, sq.data(), tl * 8))
        return -1;
    SearchParams& sp = s.sp;
    sp.nodes = s.nodes.p; sp.node_count = s.node_count.p; sp.max_nodes = max_nodes;
    sp.pos = s.pos.p; sp.pos_count = s.pos_count.p; sp.max_pos = max_positions;
    sp.trees = s.trees.p; sp.root_p64 = s.root_p64.p;
    sp.noise = s.noise.p; sp.cpuct_tab = s.cpuct_tab.p; sp.sqrt_tab = s.sqrt_tab.p; sp.tab_len = tl;
    sp.exp_table = ctx->exp_table.p; sp.counters = s.counters.p;
    s.leaves = 0;
    return cb_search_set_leaves(ctx, 1);
}

// Buffers indexed by evaluation slot (tree * leaves + i): the batch of a step is max_trees * leaves positions.
int cb_search_set_leaves(cb_ctx* ctx, int leaves) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    if (leaves < 1 || leaves > 32) { cb_set_error("leaves per tree must be 1..32"); return -1; }
    if (s.max_trees <= 0) { cb_set_error("cb_search_create first"); return -1; }
    if (leaves == s.leaves) return 0;
    const size_t ns = (size_t)s.max_trees * leaves;
    if (s.active_list.alloc(ns) || s.slot_board.alloc(ns) || s.active_count.alloc(1) || s.frame_idx.alloc(ns * 8) || s.path.alloc(ns * MCTS_MAX_PATH) || s.pend_moves.alloc(ns * MAX_MOVES) || s.value.alloc(ns) ||
        s.synth_hash.alloc(ns) || s.slots.alloc(ns) || s.planes_act.alloc(act_bytes((int)ns, 128) / 2) || s.planes_i16.alloc(ns * 7616))
        return -1;
    s.expvals.free_();
    s.leaves = leaves;
    SearchParams& sp = s.sp;
    sp.leaves = leaves; sp.slots = s.slots.p; sp.path = s.path.p; sp.pend_moves = s.pend_moves.p; sp.frame_idx = s.frame_idx.p;
    sp.value = s.value.p; sp.synth_hash = s.synth_hash.p; sp.expvals = nullptr;
    sp.active_list = s.active_list.p; sp.active_count = s.active_count.p;
    return 0;
}

// Roots + per-tree settings -> device, and the root expansion is queued.  Everything the asynchronous copies read lies in pinned
// staging owned by the context (no pageable bounce, no host synchronisation here); the staging is reused once the event recorded
// behind the copies has passed.  noise_flat / noise_off: CSR (tree t's samples at noise_flat[noise_off[t] .. noise_off[t+1]), an empty
// range = no root noise for that tree); noise_off == nullptr = no noise at all.
static int search_begin_impl(cb_ctx* ctx, int n_trees, const void* records_host, const int32_t* rec_len, const int32_t* sims, const double* cpuct,
                             const double* noise_flat, const int64_t* noise_off, int max_game_length, int eval_mode, uint32_t synth_seed,
                             cudaStream_t st) {
    Search& s = ctx->search;
    if (n_trees > s.max_trees) { cb_set_error("n_trees > max_trees"); return -1; }
    if (eval_mode == CB_EVAL_NET && !ctx->net.ready) { cb_set_error("CB_EVAL_NET without a finalized network"); return -1; }
    if (eval_mode == CB_EVAL_NET && n_trees * s.leaves > ctx->net.max_batch) { cb_set_error("n_trees * leaves > network max_batch"); return -1; }
    if (eval_mode == CB_EVAL_EXTERNAL && s.leaves != 1) { cb_set_error("a host network evaluates one leaf per tree (leaves = 1)"); return -1; }
    s.n_trees = n_trees;
    s.eval_mode = eval_mode;
    s.seed = synth_seed;
    size_t total = 0;
    for (int i = 0; i < n_trees; ++i) {
        if (rec_len[i] < 1) { cb_set_error("empty game record"); return -1; }
        if (sims[i] > s.max_sims) { cb_set_error("sims > max_sims"); return -1; }
        total += rec_len[i];
    }
    if ((int64_t)total >= s.sp.max_pos) { cb_set_error("root records exceed the position pool"); return -1; }
    const size_t n_noise = noise_off ? (size_t)noise_off[n_trees] : 0;
    if (n_noise > s.noise.n) { cb_set_error("more root noise samples than n_trees * CB_MAX_MOVES"); return -1; }
    // staging layout: positions | tree states | noise | two counters
    const size_t b_pos = total * sizeof(Pos), b_ts = (size_t)n_trees * sizeof(TreeState), b_noise = n_noise * 8;
    const size_t o_ts = (b_pos + 15) & ~(size_t)15, o_noise = (o_ts + b_ts + 15) & ~(size_t)15, o_cnt = (o_noise + b_noise + 15) & ~(size_t)15;
    if (!s.begin_copied) CB_CUDA_OK(cudaEventCreateWithFlags(&s.begin_copied, cudaEventDisableTiming));
    else CB_CUDA_OK(cudaEventSynchronize(s.begin_copied));
    if (s.h_begin.reserve(o_cnt + 16)) return -1;
    Pos* rec = reinterpret_cast<Pos*>(s.h_begin.p);
    TreeState* ts = reinterpret_cast<TreeState*>(s.h_begin.p + o_ts);
    memcpy(rec, records_host, b_pos);
    memset(ts, 0, b_ts);
    size_t off = 0;
    for (int i = 0; i < n_trees; ++i) {
        for (int k = 0; k < rec_len[i]; ++k) rec[off + k].pad_[0] = k ? (uint32_t)(off + k - 1) : 0xFFFFFFFFu;  // predecessor link
        ts[i].root_pos = (int32_t)(off + rec_len[i] - 1);
        ts[i].rec_start = (int32_t)off;
        ts[i].sims_target = sims[i];
        ts[i].status = 1;
        ts[i].noisy = noise_off && noise_off[i + 1] > noise_off[i];
        ts[i].noise_off = noise_off ? (int32_t)noise_off[i] : 0;
        ts[i].max_len = max_game_length;
        ts[i].cpuct = cpuct ? cpuct[i] : -1.0;
        off += rec_len[i];
    }
    if (b_noise) memcpy(s.h_begin.p + o_noise, noise_flat, b_noise);
    int32_t* cnt = reinterpret_cast<int32_t*>(s.h_begin.p + o_cnt);
    cnt[0] = 0;
    cnt[1] = (int32_t)total;
    CB_CUDA_OK(cudaMemcpyAsync(s.pos.p, rec, b_pos, cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaMemcpyAsync(s.trees.p, ts, b_ts, cudaMemcpyHostToDevice, st));
    if (b_noise) CB_CUDA_OK(cudaMemcpyAsync(s.noise.p, s.h_begin.p + o_noise, b_noise, cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaMemcpyAsync(s.node_count.p, &cnt[0], 4, cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaMemcpyAsync(s.pos_count.p, &cnt[1], 4, cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaEventRecord(s.begin_copied, st));
    CB_CUDA_OK(cudaMemsetAsync(s.counters.p, 0, 16, st));
    CB_CUDA_OK(cudaMemsetAsync(s.active_count.p, 0, 4, st));
    s.sp.n_trees = n_trees;
    s.sp.eval_mode = eval_mode;
    s.sp.pfeat = ctx->net.pfeat.p;
    s.sp.wpt = ctx->net.wpt.p;
    s.sp.expvals = s.expvals.p;
    // fused step with the device network: heads in mcts_step's phase A, the selected leaf encoded at the end of its phase B
    const bool fused = eval_mode == CB_EVAL_NET;
    s.sp.slot_board = s.slot_board.p;
    s.sp.head_raw = fused ? ctx->net.head_raw.p : nullptr;
    s.sp.head_bn = ctx->net.head_bn.p; s.sp.fc1 = ctx->net.fc1.p; s.sp.fc2 = ctx->net.fc2.p;
    s.sp.head_c = ctx->net.filters; s.sp.slope = ctx->net.slope;
    s.sp.planes_act = fused ? s.planes_act.p : nullptr;
    if (launch_mcts_begin(s.sp, st)) return -1;
    if (fused) {  // the roots are the first batch: encoded by the stand-alone kernel, every later batch by mcts_step itself
        if (prof_begin(ctx, CB_PROF_ENCODE, st) ||
            launch_encode(s.pos.p, s.frame_idx.p, s.n_trees * s.leaves, s.planes_act.p, nullptr, st, s.active_list.p, s.active_count.p) ||
            prof_end(ctx, st))
            return -1;
    }
    return 0;
}

int cb_search_begin(cb_ctx* ctx, int n_trees, const void* records_host, const int32_t* rec_len, const int32_t* sims,
                    const double* cpuct, const uint8_t* has_noise, const double* noise, int max_game_length, int eval_mode,
                    uint32_t synth_seed, void* stream) {
    CB_ENTER(ctx);
    if (!noise || !has_noise)
        return search_begin_impl(ctx, n_trees, records_host, rec_len, sims, cpuct, nullptr, nullptr, max_game_length, eval_mode, synth_seed, (cudaStream_t)stream);
    // dense [n][CB_MAX_MOVES] rows -> CSR with full-length rows for the noisy trees
    std::vector<int64_t> off(n_trees + 1, 0);
    std::vector<double> flat;
    flat.reserve((size_t)n_trees * MAX_MOVES);
    for (int i = 0; i < n_trees; ++i) {
        if (has_noise[i]) flat.insert(flat.end(), noise + (size_t)i * MAX_MOVES, noise + (size_t)(i + 1) * MAX_MOVES);
        off[i + 1] = (int64_t)flat.size();
    }
    return search_begin_impl(ctx, n_trees, records_host, rec_len, sims, cpuct, flat.data(), off.data(), max_game_length, eval_mode, synth_seed,
                             (cudaStream_t)stream);
}
int cb_search_begin_csr(cb_ctx* ctx, int n_trees, const void* records_host, const int32_t* rec_len, const int32_t* sims, const double* cpuct,
                        const double* noise_flat, const int64_t* noise_off, int max_game_length, int eval_mode, uint32_t synth_seed, void* stream) {
    CB_ENTER(ctx);
    return search_begin_impl(ctx, n_trees, records_host, rec_len, sims, cpuct, noise_flat, noise_flat ? noise_off : nullptr, max_game_length, eval_mode,
                             synth_seed, (cudaStream_t)stream);
}

// CB_DEBUG_SYNC=1: synchronise after every stage of a search step and name the stage that faulted
static int debug_sync(cudaStream_t st, const char* stage) {
    static const bool on = getenv("CB_DEBUG_SYNC") != nullptr;
    if (!on) return 0;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cb_set_error(std::string("CB_DEBUG_SYNC: ") + stage + ": " + cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

int cb_search_step(cb_ctx* ctx, void* stream) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    cudaStream_t st = (cudaStream_t)stream;
    if (s.eval_mode == CB_EVAL_NET) {
        // only the slots that hold a pending leaf are in the batch (compacted, size known on the device only); their planes were written by
        // the previous mcts_step (the roots' by cb_search_begin), and mcts_step finishes the heads: two launches per step
        if (net_tower(ctx, s.planes_act.p, s.n_trees * s.leaves, s.value.p, st, s.active_list.p, s.active_count.p, /*run_heads=*/false) ||
            debug_sync(st, "tower"))
            return -1;
    } else if (s.eval_mode == CB_EVAL_SYNTH) {
        if (launch_encode(s.pos.p, s.frame_idx.p, s.n_trees * s.leaves, nullptr, s.planes_i16.p, st)) return -1;
        if (launch_synth(s.planes_i16.p, s.n_trees * s.leaves, s.seed, s.synth_hash.p, s.value.p, nullptr, st)) return -1;
    } else {
        cb_set_error("cb_search_step: external evaluation uses cb_search_step_external");
        return -1;
    }
    CB_CUDA_OK(cudaMemsetAsync(s.active_count.p, 0, 4, st));  // mcts_step claims the next step's batch
    if (prof_begin(ctx, CB_PROF_MCTS, st) || launch_mcts_step(s.sp, st) || prof_end(ctx, st)) return -1;
    return debug_sync(st, "mcts_step");
}
int cb_search_run(cb_ctx* ctx, int steps, void* stream) {
    for (int i = 0; i < steps; ++i)
        if (cb_search_step(ctx, stream)) return -1;
    return 0;
}
int cb_search_pending_planes_sync(cb_ctx* ctx, int16_t* planes_host, uint8_t* waiting_host, void* stream) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    cudaStream_t st = (cudaStream_t)stream;
    if (launch_encode(s.pos.p, s.frame_idx.p, s.n_trees, nullptr, s.planes_i16.p, st)) return -1;
    waiting_flags_kernel<<<(s.n_trees + 127) / 128, 128, 0, st>>>(s.trees.p, s.n_trees, s.flags.p);
    CB_CUDA_OK(cudaMemcpyAsync(planes_host, s.planes_i16.p, (size_t)s.n_trees * 7616 * 2, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(waiting_host, s.flags.p, s.n_trees, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
int cb_search_step_external(cb_ctx* ctx, const float* value_host, const float* expvals_host, void* stream) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    cudaStream_t st = (cudaStream_t)stream;
    if (s.expvals.n < (size_t)s.n_trees * 4672) {
        if (s.expvals.alloc((size_t)s.max_trees * 4672)) return -1;
        s.sp.expvals = s.expvals.p;
    }
    CB_CUDA_OK(cudaMemcpyAsync(s.value.p, value_host, (size_t)s.n_trees * 4, cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaMemcpyAsync(s.expvals.p, expvals_host, (size_t)s.n_trees * 4672 * 4, cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaMemsetAsync(s.active_count.p, 0, 4, st));
    if (launch_mcts_step(s.sp, st)) return -1;
    CB_CUDA_OK(cudaStreamSynchronize(st));  // the host buffers may be reused by the caller
    return 0;
}
int cb_search_status_sync(cb_ctx* ctx, int64_t* counters, void* stream) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA_OK(cudaMemsetAsync(s.scratch_i32.p, 0, 4, st));
    count_unfinished_kernel<<<(s.n_trees + 127) / 128, 128, 0, st>>>(s.trees.p, s.n_trees, s.scratch_i32.p);
    int32_t unfinished = 0, c[4];
    CB_CUDA_OK(cudaMemcpyAsync(&unfinished, s.scratch_i32.p, 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(c, s.counters.p, 16, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));
    counters[0] = unfinished; counters[1] = c[0]; counters[2] = c[1]; counters[3] = c[2];
    return 0;
}
int cb_search_read_roots_sync(cb_ctx* ctx, uint16_t* moves, int32_t* n, float* w, float* q, double* p, int32_t* n_children,
                              int32_t* root_n, int32_t* tree_nodes, void* stream) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t mm = (size_t)s.n_trees * MAX_MOVES, mt = s.n_trees;
    read_roots_kernel<<<s.n_trees, 64, 0, st>>>(s.sp, s.r_moves.p, s.r_n.p, s.r_w.p, s.r_q.p, s.r_p.p, s.r_nc.p, s.r_rootn.p, s.r_nodes.p);
    CB_CUDA_OK(cudaGetLastError());
    CB_CUDA_OK(cudaMemcpyAsync(moves, s.r_moves.p, mm * 2, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(n, s.r_n.p, mm * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(w, s.r_w.p, mm * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(q, s.r_q.p, mm * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(p, s.r_p.p, mm * 8, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(n_children, s.r_nc.p, mt * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(root_n, s.r_rootn.p, mt * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(tree_nodes, s.r_nodes.p, mt * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

size_t cb_search_roots_compact_bytes(int n_trees, int total_children) {
    return (size_t)total_children * (8 + 4 + 4 + 4) + (size_t)n_trees * 16 + (size_t)total_children * 2;
}
// Root statistics of all trees in ONE packed buffer and ONE device-to-host copy: the children of tree t at [child_off[t], child_off[t+1])
// of every section (the caller knows every root's legal-move count from the record export).  Sections, in this order:
// P f64[T] | N i32[T] | W f32[T] | Q f32[T] | n_children i32[n] | root_N i32[n] | tree_nodes i32[n] | state i32[n] | moves u16[T],  T = child_off[n].
// _async queues kernel + copy + an event behind the search steps already on the stream and returns; _wait blocks on that event only,
// so the host can pick up one search's results while the next one (another context, same stream) already runs.
int cb_search_read_roots_compact_async(cb_ctx* ctx, const int32_t* child_off, void* stream) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = s.n_trees, total = child_off[n];
    if (n < 1) { cb_set_error("read_roots_compact: no search"); return -1; }
    if (total < 0 || (size_t)total > (size_t)s.max_trees * MAX_MOVES) { cb_set_error("read_roots_compact: bad offsets"); return -1; }
    const size_t bytes = cb_search_roots_compact_bytes(n, total), b_off = (size_t)(n + 1) * 4;
    if (s.r_pack.n < bytes && s.r_pack.alloc(cb_search_roots_compact_bytes(s.max_trees, s.max_trees * MAX_MOVES), false)) return -1;
    if (s.r_off.n < (size_t)n + 1 && s.r_off.alloc(s.max_trees + 1, false)) return -1;
    const size_t o_cnt = (bytes + 15) & ~(size_t)15, o_off = o_cnt + 16;
    if (!s.read_done) CB_CUDA_OK(cudaEventCreateWithFlags(&s.read_done, cudaEventDisableTiming));
    else CB_CUDA_OK(cudaEventSynchronize(s.read_done));   // an earlier read-back nobody waited for still owns the staging
    if (s.h_read.reserve(o_off + b_off)) return -1;
    memcpy(s.h_read.p + o_off, child_off, b_off);
    CB_CUDA_OK(cudaMemcpyAsync(s.r_off.p, s.h_read.p + o_off, b_off, cudaMemcpyHostToDevice, st));
    read_roots_compact_kernel<<<n, 64, 0, st>>>(s.sp, s.r_off.p, s.r_pack.p, total);
    CB_CUDA_OK(cudaGetLastError());
    CB_CUDA_OK(cudaMemcpyAsync(s.h_read.p, s.r_pack.p, bytes, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaMemcpyAsync(s.h_read.p + o_cnt, s.counters.p, 16, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaEventRecord(s.read_done, st));
    s.read_bytes = bytes;
    s.read_cnt_off = o_cnt;
    return 0;
}
// counters[4] like cb_search_status_sync: trees not finished, evaluations, simulations, errors
int cb_search_read_roots_compact_wait(cb_ctx* ctx, void* out_host, int64_t* counters) {
    CB_ENTER(ctx);
    Search& s = ctx->search;
    if (!s.read_done || !s.read_bytes) { cb_set_error("read_roots_compact_wait: nothing queued"); return -1; }
    CB_CUDA_OK(cudaEventSynchronize(s.read_done));
    memcpy(out_host, s.h_read.p, s.read_bytes);
    if (counters) {
        const int32_t* c = reinterpret_cast<const int32_t*>(s.h_read.p + s.read_cnt_off);
        const int n = s.n_trees;
        // the state section sits behind P | N | W | Q and three header rows: recover T from the byte count
        const size_t T = (s.read_bytes - (size_t)n * 16) / 22;
        const int32_t* state = reinterpret_cast<const int32_t*>(s.h_read.p + T * 20) + 3 * (size_t)n;
        int64_t unfinished = 0, errors = c[2];
        for (int i = 0; i < n; ++i) {
            const int stt = state[i] & 255;
            if (stt != TREE_DONE && stt != TREE_IDLE) ++unfinished;
            if (state[i] & 256) errors = errors ? errors : 1;
        }
        counters[0] = unfinished; counters[1] = c[0]; counters[2] = c[1]; counters[3] = errors;
    }
    s.read_bytes = 0;
    return 0;
}
int cb_search_read_roots_compact_sync(cb_ctx* ctx, const int32_t* child_off, void* out_host, void* stream) {
    if (cb_search_read_roots_compact_async(ctx, child_off, stream)) return -1;
    return cb_search_read_roots_compact_wait(ctx, out_host, nullptr);
}

}  // extern "C"
