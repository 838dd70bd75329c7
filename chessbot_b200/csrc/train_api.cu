// Device side of the training-step C ABI (include/chessbot_b200.h, cb_train_*): one SGD step of train.py:150-164 /
// model.py:59-70 on a batch of replay samples.  Parameters, gradients and momentum live in three caller-owned flat f32
// device buffers (PyTorch allocates them, and all-reduces the gradient buffer over NCCL between cb_train_forward_backward
// and cb_train_apply when several GPUs train data-parallel); everything else is owned here.
//
// Per convolution layer the step runs: forward  = tower kernel (raw convolution, bf16) -> batch statistics -> BN + skip + LeakyReLU;
// backward = BN backward (two passes) -> wgrad.cu -> tower kernel on the flipped / transposed weights (data gradient, + the
// residual branch's gradient as its "skip").  The raw convolution Y_l and the activation A_l of every layer stay in HBM for
// the backward pass (2 x 41 x 210 MB at 20x256, batch 4096: 17 GB of the 180 GB).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/chessbot_b200.h"
#include "kernels.cuh"

using namespace cb;

namespace {
template <class T>
struct TBuf {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        if (p) cb_device_free(p);
        p = nullptr;
        n = count;
        if (!count) return 0;
        if (cb_device_alloc((void**)&p, count * sizeof(T))) return -1;
        CB_CUDA_OK(cudaMemset(p, 0, count * sizeof(T)));
        return 0;
    }
    ~TBuf() { if (p) cb_device_free(p); }
    TBuf() = default;
    TBuf(TBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    TBuf(const TBuf&) = delete;
    TBuf& operator=(const TBuf&) = delete;
};

struct Span { size_t off = 0, count = 0; };

struct ParamLayout {
    int C = 0, blocks = 0, nl = 0;
    std::vector<Span> w, bn;     // per conv layer: kernel; gamma|beta
    Span vconv, vfc1, vfc2, pconv, pfc, vbn, pbn;
    std::vector<Span> mov;       // per conv layer: moving mean|var (separate buffer)
    Span vmov, pmov;
    size_t n_params = 0, n_reg = 0, n_moving = 0;
    ParamLayout(int filters, int blocks_) : C(filters), blocks(blocks_), nl(1 + 2 * blocks_) {
        size_t o = 0;
        auto take = [&](size_t cnt) { Span s{o, cnt}; o += cnt; return s; };
        w.resize(nl); bn.resize(nl); mov.resize(nl);
        for (int l = 0; l < nl; ++l) w[l] = take((size_t)9 * (l == 0 ? 119 : C) * C);
        vconv = take(C); vfc1 = take((size_t)64 * C); vfc2 = take(C); pconv = take((size_t)2 * C); pfc = take((size_t)128 * 4672);
        n_reg = o;
        for (int l = 0; l < nl; ++l) bn[l] = take((size_t)2 * C);
        vbn = take(2); pbn = take(4);
        n_params = o;
        size_t m = 0;
        for (int l = 0; l < nl; ++l) { mov[l] = Span{m, (size_t)2 * C}; m += 2 * C; }
        vmov = Span{m, 2}; m += 2;
        pmov = Span{m, 4}; m += 4;
        n_moving = m;
    }
};

enum { TP_CONV_FWD = 0, TP_BN_FWD, TP_HEADS, TP_BN_BWD, TP_WGRAD, TP_DGRAD, TP_OPT, TP_INPUT, TP_STAGES };
}  // namespace

struct cb_trainer {
    int device = 0, sms = 0, C = 0, N = 0, blocks = 0, nl = 0, batch = 0, pairs = 0;
    float slope = 0.3f, eps = 1e-3f, bn_momentum = 0.99f, last_l2 = 0.f;
    ParamLayout lay;
    float *params = nullptr, *grads = nullptr, *mom = nullptr;
    TBuf<float> moving;
    TBuf<__nv_bfloat16> planes, g[2], dy[2], dz;  // dy: double-buffered, the weight gradient of layer l reads dY_l while layer l - 1 is at work
    std::vector<TBuf<__nv_bfloat16>> Y, A, wf, wt;
    std::vector<CUtensorMap> map_A;      // A[l] as conv / wgrad input (box 200 rows)
    CUtensorMap map_planes{}, map_dy_conv[2]{}, map_dy_wg[2]{};
    TBuf<CUtensorMap> wmaps;             // [2 * nl]: forward packs, then data-gradient packs
    TBuf<LayerDesc> ldesc;               // [2 * nl]
    TBuf<float> ones, zeros, mean_rstd, wg_partial, gemm_scratch;
    TBuf<double> stats, bsums, hstats, hsums, losses;
    TBuf<int> flags;
    int wg_splits_stem = 1, wg_splits_res = 1;
    TBuf<float> head_raw, hmr, vfeat, pfeat, h1pre, value, logits, dpfeat, dvfeat, dh1pre, dpre;
    // the step's launch sequence is the same every time: after one eager step it is captured into two CUDA graphs (forward +
    // backward; optimizer + weight packing) and replayed.  Inputs are staged into trainer-owned buffers, the optimizer's
    // hyper-parameters live in device memory, so no kernel argument changes between steps.
    TBuf<int16_t> in_img;
    TBuf<float> in_z, in_pi, hp;
    cudaStream_t cap_stream = nullptr, aux_stream = nullptr;   // aux: the weight-gradient branch of the backward pass
    std::vector<cudaEvent_t> ev_ready, ev_wdone;               // per layer: dY_l final and the data gradient done; weight gradient done
    cudaGraphExec_t graph_fb = nullptr, graph_apply = nullptr;
    int fb_calls = 0, apply_calls = 0;
    bool weights_dirty = true;
    bool profile = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_stage;
    size_t ev_used = 0;
    cb_trainer(int filters, int blocks_) : lay(filters, blocks_) {}
};

static int tp_begin(cb_trainer* t, int stage, cudaStream_t st) {
    if (!t->profile) return 0;
    while (t->ev.size() < t->ev_used + 2) {
        cudaEvent_t e;
        CB_CUDA_OK(cudaEventCreate(&e));
        t->ev.push_back(e);
    }
    if (t->ev_stage.size() < t->ev_used / 2 + 1) t->ev_stage.resize(t->ev_used / 2 + 1);
    t->ev_stage[t->ev_used / 2] = stage;
    CB_CUDA_OK(cudaEventRecord(t->ev[t->ev_used], st));
    return 0;
}
static int tp_end(cb_trainer* t, cudaStream_t st) {
    if (!t->profile) return 0;
    CB_CUDA_OK(cudaEventRecord(t->ev[t->ev_used + 1], st));
    t->ev_used += 2;
    return 0;
}

// one layer of the tower kernel: raw convolution (identity epilogue), optionally + skip
static int conv_layer(cb_trainer* t, int desc, const CUtensorMap& in_map, const void* in, void* out, const void* skip, double* stats,
                      cudaStream_t st) {
    TowerParams tp{};  // one layer per launch: nothing waits on the progress flags, so the epoch can stay the same (graph replay)
    tp.tm_act[0] = in_map;
    tp.bufs[0] = (uint8_t*)const_cast<void*>(in);
    tp.bufs[1] = (uint8_t*)out;
    tp.bufs[2] = (uint8_t*)const_cast<void*>(skip);
    tp.maps = t->wmaps.p;
    tp.layers = t->ldesc.p + desc;
    tp.head_w = nullptr;
    tp.head_out = nullptr;
    tp.stats = stats;  // forward pass: the epilogue gathers the layer's batch statistics
    tp.full_sectors = act_bytes(t->batch, t->N) > ((size_t)32 << 20);  // a layer's tensors are read back after ~10 other tensors were streamed
    tp.flags = t->flags.p;
    tp.epoch = 1;
    tp.nl = 1;
    tp.pairs = t->pairs;
    tp.n = t->N;
    tp.slope = 1.0f;  // identity: BatchNormalization and LeakyReLU follow once the batch statistics are known
    return launch_tower(tp, t->sms, st);
}

static int pack_weights(cb_trainer* t, cudaStream_t st) {
    for (int l = 0; l < t->nl; ++l)
        if (launch_pack_conv(t->params + t->lay.w[l].off, l == 0 ? 119 : t->C, l == 0 ? 128 : t->N, t->C, t->N, l == 0, t->wf[l].p,
                             l == 0 ? nullptr : t->wt[l].p, st))
            return -1;
    t->weights_dirty = false;
    return 0;
}

extern "C" {

size_t cb_train_param_count(int filters, int blocks, size_t* n_regularized) {
    ParamLayout lay(filters, blocks);
    if (n_regularized) *n_regularized = lay.n_reg;
    return lay.n_params;
}

cb_trainer* cb_train_create(cb_ctx* ctx, int filters, int blocks, int batch, float leaky_slope, float bn_eps, float bn_momentum, float* params_dev,
                            float* grads_dev, float* momentum_dev) {
    if (!ctx) { cb_set_error("cb_train_create: no context"); return nullptr; }
    cb_device_guard cb_guard_(cb_ctx_device(ctx));
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(context device) failed"); return nullptr; }
    const int N = (filters + 63) / 64 * 64;
    if (filters < 4 || (filters & 3) || (N != 64 && N != 128 && N != 256) || blocks < 0 || batch < 1) {
        cb_set_error("cb_train_create: filters must be a multiple of 4 that pads to 64, 128 or 256 channels (4..64, 68..128, 196..256), batch >= 1");
        return nullptr;
    }
    cb_trainer* t = new cb_trainer(filters, blocks);
    t->sms = cb_ctx_sm_count(ctx);
    t->device = cb_ctx_device(ctx);
    t->C = filters; t->N = N; t->blocks = blocks; t->nl = 1 + 2 * blocks; t->batch = batch; t->pairs = (batch + 1) / 2;
    t->slope = leaky_slope; t->eps = bn_eps; t->bn_momentum = bn_momentum;
    t->params = params_dev; t->grads = grads_dev; t->mom = momentum_dev;
    const int nl = t->nl, cj = N / 8;
    const size_t act_elems = act_bytes(batch, N) / 2;
    auto fail = [&]() { delete t; return (cb_trainer*)nullptr; };
    t->Y.resize(nl); t->A.resize(nl); t->wf.resize(nl); t->wt.resize(nl); t->map_A.resize(nl);
    if (t->planes.alloc(act_bytes(batch, 128) / 2) || t->g[0].alloc(act_elems) || t->g[1].alloc(act_elems) || t->dy[0].alloc(act_elems) || t->dy[1].alloc(act_elems) ||
        t->dz.alloc(act_elems) || t->moving.alloc(t->lay.n_moving))
        return fail();
    for (int l = 0; l < nl; ++l) {
        const int kc = l == 0 ? 4 : N / 64;  // stem: 128 input channels x (hi, lo)
        if (t->Y[l].alloc(act_elems) || t->A[l].alloc(act_elems) || t->wf[l].alloc((size_t)kc * 9 * N * 64) ||
            (l > 0 && t->wt[l].alloc((size_t)(N / 64) * 9 * N * 64)))
            return fail();
        if (conv_make_row_map(&t->map_A[l], t->A[l].p, (uint64_t)t->pairs * cj * 25, 200)) return fail();
    }
    const int dy_chunks = N / 16;
    if (conv_make_row_map(&t->map_planes, t->planes.p, (uint64_t)t->pairs * 16 * 25, 200) ||
        conv_make_row_map(&t->map_dy_conv[0], t->dy[0].p, (uint64_t)t->pairs * cj * 25, 200) ||
        conv_make_row_map(&t->map_dy_conv[1], t->dy[1].p, (uint64_t)t->pairs * cj * 25, 200) ||
        conv_make_row_map(&t->map_dy_wg[0], t->dy[0].p, (uint64_t)t->pairs * cj * 25, 25 * (dy_chunks < 8 ? dy_chunks : 8)) ||
        conv_make_row_map(&t->map_dy_wg[1], t->dy[1].p, (uint64_t)t->pairs * cj * 25, 25 * (dy_chunks < 8 ? dy_chunks : 8)))
        return fail();
    // weight tensor maps and one-layer plans: [0, nl) forward, [nl, 2 nl) data gradient
    std::vector<CUtensorMap> wm(2 * nl);
    std::vector<LayerDesc> ld(2 * nl);
    if (t->ones.alloc(256) || t->zeros.alloc(256) || launch_fill(t->ones.p, 256, 1.f, nullptr)) return fail();
    const uint32_t box = conv_taps_per_box(N) * N / 2;
    for (int l = 0; l < nl; ++l) {
        const int kc = l == 0 ? 4 : N / 64;
        if (conv_make_row_map(&wm[l], t->wf[l].p, (uint64_t)kc * 9 * N, box)) return fail();
        LayerDesc& d = ld[l];
        memset(&d, 0, sizeof(d));
        d.scale = t->ones.p; d.shift = t->zeros.p;
        d.kc = kc; d.kc_in = l == 0 ? 2 : N / 64; d.cj_in = l == 0 ? 16 : cj;
        d.in_buf = 0; d.out_buf = 1; d.skip_buf = -1; d.w_map = l;
        LayerDesc& b = ld[nl + l];
        b = d;
        if (l > 0) {
            if (conv_make_row_map(&wm[nl + l], t->wt[l].p, (uint64_t)(N / 64) * 9 * N, box)) return fail();
            b.kc = N / 64; b.kc_in = N / 64; b.cj_in = cj;
            b.skip_buf = (l & 1) ? 2 : -1;  // the first convolution of a block: + the gradient of the block's residual branch
            b.w_map = nl + l;
        } else {
            wm[nl] = wm[0];
        }
    }
    if (t->wmaps.alloc(2 * nl) || t->ldesc.alloc(2 * nl) ||
        cudaMemcpy(t->wmaps.p, wm.data(), wm.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t->ldesc.p, ld.data(), ld.size() * sizeof(LayerDesc), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail();
    t->wg_splits_stem = wgrad_splits(t->sms, 16);
    t->wg_splits_res = wgrad_splits(t->sms, cj);
    const size_t part_stem = (size_t)t->wg_splits_stem * 9 * 128 * N, part_res = (size_t)t->wg_splits_res * 9 * N * N;
    const size_t mb = batch;
    if (t->wg_partial.alloc(part_stem > part_res ? part_stem : part_res) || t->stats.alloc((size_t)nl * 2 * N) ||
        t->mean_rstd.alloc((size_t)nl * 4 * N) || t->bsums.alloc((size_t)nl * 2 * N) || t->hstats.alloc(6) || t->hsums.alloc(6) || t->losses.alloc(4) ||
        t->flags.alloc(2 * ((size_t)batch / 4 + 2)) || t->head_raw.alloc(mb * 192) || t->hmr.alloc(6) || t->vfeat.alloc(mb * 64) ||
        t->pfeat.alloc(mb * 128) || t->h1pre.alloc(mb * filters) || t->value.alloc(mb) || t->logits.alloc(mb * 4672) || t->dpfeat.alloc(mb * 128) ||
        t->dvfeat.alloc(mb * 64) || t->dh1pre.alloc(mb * filters) || t->dpre.alloc(mb) || t->gemm_scratch.alloc((size_t)8 << 20) ||
        t->in_img.alloc(mb * 7616) || t->in_z.alloc(mb) || t->in_pi.alloc(mb * 4672) || t->hp.alloc(8))
        return fail();
    // Keras initial state of the moving statistics: mean 0, variance 1
    std::vector<float> mv(t->lay.n_moving, 0.f);
    for (int l = 0; l < nl; ++l)
        for (int c = 0; c < filters; ++c) mv[t->lay.mov[l].off + filters + c] = 1.f;
    mv[t->lay.vmov.off + 1] = 1.f;
    mv[t->lay.pmov.off + 2] = mv[t->lay.pmov.off + 3] = 1.f;
    if (cudaMemcpy(t->moving.p, mv.data(), mv.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { cb_set_error("moving statistics upload failed"); return fail(); }
    if (cudaStreamCreateWithFlags(&t->cap_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&t->aux_stream, cudaStreamNonBlocking) != cudaSuccess) { cb_set_error("cb_train_create: stream"); return fail(); }
    t->ev_ready.resize(nl); t->ev_wdone.resize(nl);
    for (int l = 0; l < nl; ++l)
        if (cudaEventCreateWithFlags(&t->ev_ready[l], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&t->ev_wdone[l], cudaEventDisableTiming) != cudaSuccess) { cb_set_error("cb_train_create: event"); return fail(); }
    if (cudaDeviceSynchronize() != cudaSuccess) { cb_set_error("cb_train_create: device error"); return fail(); }
    return t;
}

void cb_train_destroy(cb_trainer* t) {
    if (!t) return;
    cb_device_guard cb_guard_(t->device);
    if (!t) return;
    for (cudaEvent_t e : t->ev) cudaEventDestroy(e);
    if (t->graph_fb) cudaGraphExecDestroy(t->graph_fb);
    if (t->graph_apply) cudaGraphExecDestroy(t->graph_apply);
    if (t->cap_stream) cudaStreamDestroy(t->cap_stream);
    if (t->aux_stream) cudaStreamDestroy(t->aux_stream);
    for (cudaEvent_t e : t->ev_ready) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : t->ev_wdone) if (e) cudaEventDestroy(e);
    delete t;
}

// name -> (span in params, span in moving, channel count of a BN tensor)
static int find_tensor(cb_trainer* t, const std::string& name, Span* par, Span* mov, int* bn_c) {
    const ParamLayout& L = t->lay;
    *bn_c = 0;
    auto conv_index = [&](const std::string& base, const std::string& suffix) -> int {
        if (base == "stem") return suffix == "w" || suffix == "bn" ? 0 : -1;
        if (base.rfind("res", 0) != 0) return -1;
        const int i = atoi(base.c_str() + 3);
        if (i < 0 || i >= L.blocks) return -1;
        if (suffix == "w1" || suffix == "bn1") return 1 + 2 * i;
        if (suffix == "w2" || suffix == "bn2") return 2 + 2 * i;
        return -1;
    };
    const size_t dot = name.find('.');
    if (dot == std::string::npos) return -1;
    const std::string base = name.substr(0, dot), suffix = name.substr(dot + 1);
    if (base == "value") {
        if (suffix == "conv") { *par = L.vconv; return 0; }
        if (suffix == "fc1") { *par = L.vfc1; return 0; }
        if (suffix == "fc2") { *par = L.vfc2; return 0; }
        if (suffix == "bn") { *par = L.vbn; *mov = L.vmov; *bn_c = 1; return 0; }
        return -1;
    }
    if (base == "policy") {
        if (suffix == "conv") { *par = L.pconv; return 0; }
        if (suffix == "fc") { *par = L.pfc; return 0; }
        if (suffix == "bn") { *par = L.pbn; *mov = L.pmov; *bn_c = 2; return 0; }
        return -1;
    }
    const int l = conv_index(base, suffix);
    if (l < 0) return -1;
    if (suffix[0] == 'w') { *par = L.w[l]; return 0; }
    *par = L.bn[l]; *mov = L.mov[l]; *bn_c = L.C;
    return 0;
}

// what = 0: parameter (BN: gamma, beta, moving mean, moving variance), 2: momentum of the optimizer (BN: rows 2, 3 ignored)
int cb_train_set_tensor(cb_trainer* t, const char* name, int what, const float* host, size_t count, void* stream) {
    cb_device_guard cb_guard_(t->device);
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(trainer device) failed"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    Span par, mov;
    int bn_c;
    if (what != 0 && what != 2) { cb_set_error("cb_train_set_tensor: what must be 0 (parameter) or 2 (momentum)"); return -1; }
    if (find_tensor(t, name, &par, &mov, &bn_c)) { cb_set_error(std::string("unknown tensor: ") + name); return -1; }
    if (count != (bn_c ? (size_t)4 * bn_c : par.count)) { cb_set_error(std::string("tensor has wrong size: ") + name); return -1; }
    CB_CUDA_OK(cudaMemcpyAsync((what == 0 ? t->params : t->mom) + par.off, host, par.count * 4, cudaMemcpyHostToDevice, st));
    if (bn_c && what == 0) CB_CUDA_OK(cudaMemcpyAsync(t->moving.p + mov.off, host + 2 * bn_c, mov.count * 4, cudaMemcpyHostToDevice, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));
    if (what == 0) t->weights_dirty = true;
    return 0;
}

// what = 0: parameter (BN: gamma, beta, moving mean, moving variance), 1: gradient (BN: dgamma, dbeta, 0, 0), 2: momentum
int cb_train_get_tensor(cb_trainer* t, const char* name, int what, float* host, size_t count, void* stream) {
    cb_device_guard cb_guard_(t->device);
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(trainer device) failed"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    Span par, mov;
    int bn_c;
    if (find_tensor(t, name, &par, &mov, &bn_c)) { cb_set_error(std::string("unknown tensor: ") + name); return -1; }
    if (count != (bn_c ? (size_t)4 * bn_c : par.count)) { cb_set_error(std::string("tensor has wrong size: ") + name); return -1; }
    const float* src = what == 0 ? t->params : (what == 1 ? t->grads : t->mom);
    CB_CUDA_OK(cudaMemcpyAsync(host, src + par.off, par.count * 4, cudaMemcpyDeviceToHost, st));
    if (bn_c) {
        if (what == 0) CB_CUDA_OK(cudaMemcpyAsync(host + 2 * bn_c, t->moving.p + mov.off, mov.count * 4, cudaMemcpyDeviceToHost, st));
        else memset(host + 2 * bn_c, 0, (size_t)2 * bn_c * 4);
    }
    CB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

static int enqueue_forward_backward(cb_trainer* t, cudaStream_t st);
static int enqueue_apply(cb_trainer* t, cudaStream_t st);

// runs `enqueue` eagerly the first time (and whenever per-launch profiling is on or CB_TRAIN_NO_GRAPH is set), captures it into a graph
// on the trainer's own stream the second time (the caller's stream may be the legacy default stream, which cannot capture) and
// replays that graph on the caller's stream from then on
static int run_graphed(cb_trainer* t, int (*enqueue)(cb_trainer*, cudaStream_t), cudaGraphExec_t* exec, int* calls, cudaStream_t st) {
    static const bool no_graph = getenv("CB_TRAIN_NO_GRAPH") != nullptr;
    if (t->profile || no_graph || (*calls)++ == 0) return enqueue(t, st);
    if (!*exec) {
        cudaGraph_t graph = nullptr;
        CB_CUDA_OK(cudaStreamBeginCapture(t->cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue(t, t->cap_stream);
        const cudaError_t e = cudaStreamEndCapture(t->cap_stream, &graph);
        if (rc || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            if (!rc) cb_set_error(std::string("CUDA graph capture of the training step failed: ") + cudaGetErrorString(e));
            return -1;
        }
        const cudaError_t ei = cudaGraphInstantiate(exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { cb_set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei)); return -1; }
    }
    CB_CUDA_OK(cudaGraphLaunch(*exec, st));
    return 0;
}

int cb_train_forward_backward(cb_trainer* t, const int16_t* images_dev, const float* z_dev, const float* pi_dev, int n, void* stream) {
    cb_device_guard cb_guard_(t->device);
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(trainer device) failed"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n != t->batch) { cb_set_error("cb_train_forward_backward: n must equal the trainer's batch"); return -1; }
    if (t->weights_dirty && (tp_begin(t, TP_OPT, st) || pack_weights(t, st) || tp_end(t, st))) return -1;
    // stage the caller's batch: the step itself only ever reads the trainer's own buffers
    CB_CUDA_OK(cudaMemcpyAsync(t->in_img.p, images_dev, (size_t)n * 7616 * 2, cudaMemcpyDeviceToDevice, st));
    CB_CUDA_OK(cudaMemcpyAsync(t->in_z.p, z_dev, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    CB_CUDA_OK(cudaMemcpyAsync(t->in_pi.p, pi_dev, (size_t)n * 4672 * 4, cudaMemcpyDeviceToDevice, st));
    return run_graphed(t, enqueue_forward_backward, &t->graph_fb, &t->fb_calls, st);
}

static int enqueue_forward_backward(cb_trainer* t, cudaStream_t st) {
    const int n = t->batch;
    const int16_t* images_dev = t->in_img.p;
    const float* z_dev = t->in_z.p;
    const float* pi_dev = t->in_pi.p;
    const int nl = t->nl, N = t->N, C = t->C, cj = N / 8;
    const ParamLayout& L = t->lay;
    float* P = t->params;
    float* G = t->grads;
    CB_CUDA_OK(cudaMemsetAsync(t->stats.p, 0, t->stats.n * 8, st));
    CB_CUDA_OK(cudaMemsetAsync(t->bsums.p, 0, t->bsums.n * 8, st));
    CB_CUDA_OK(cudaMemsetAsync(t->hstats.p, 0, 48, st));
    CB_CUDA_OK(cudaMemsetAsync(t->hsums.p, 0, 48, st));
    CB_CUDA_OK(cudaMemsetAsync(t->losses.p, 0, 16, st));
    if (tp_begin(t, TP_INPUT, st) || launch_planes_from_images(images_dev, n, t->planes.p, st) || tp_end(t, st)) return -1;

    // ---------------------------------------------------------------- forward tower
    for (int l = 0; l < nl; ++l) {
        const void* in = l == 0 ? (const void*)t->planes.p : (const void*)t->A[l - 1].p;
        // CB_TRAIN_SEPARATE_STATS=1 (measurement only): batch statistics by a separate pass over Y instead of the conv epilogue
        static const bool separate_stats = getenv("CB_TRAIN_SEPARATE_STATS") != nullptr;
        double* stats = t->stats.p + (size_t)l * 2 * N;
        if (tp_begin(t, TP_CONV_FWD, st) ||
            conv_layer(t, l, l == 0 ? t->map_planes : t->map_A[l - 1], in, t->Y[l].p, nullptr, separate_stats ? nullptr : stats, st) || tp_end(t, st))
            return -1;
        const __nv_bfloat16* skip = (l > 0 && (l & 1) == 0) ? t->A[l - 2].p : nullptr;
        if (tp_begin(t, TP_BN_FWD, st) || (separate_stats && launch_bn_stats(t->Y[l].p, t->pairs, cj, stats, st)) ||
            launch_bn_apply(t->Y[l].p, skip, t->A[l].p, n, cj, C, stats, P + L.bn[l].off, P + L.bn[l].off + C, t->mean_rstd.p + (size_t)l * 4 * N,
                            t->moving.p + L.mov[l].off, t->eps, t->slope, t->bn_momentum, st) ||
            tp_end(t, st))
            return -1;
    }
    // ---------------------------------------------------------------- heads, losses, head gradients
    const __nv_bfloat16* top = t->A[nl - 1].p;
    if (tp_begin(t, TP_HEADS, st)) return -1;
    if (launch_head_conv_fwd(top, n, cj, P + L.vconv.off, P + L.pconv.off, C, t->head_raw.p, st) || launch_head_stats(t->head_raw.p, n, t->hstats.p, st) ||
        launch_head_features_train(t->head_raw.p, n, t->hstats.p, P + L.vbn.off, P + L.pbn.off, t->moving.p + L.vmov.off, t->moving.p + L.pmov.off,
                                   P + L.vfc1.off, P + L.vfc2.off, C, t->eps, t->slope, t->bn_momentum, t->hmr.p, t->vfeat.p, t->pfeat.p, t->h1pre.p,
                                   t->value.p, st) ||
        launch_sgemm(t->pfeat.p, P + L.pfc.off, t->logits.p, n, 4672, 128, 0, 0, t->gemm_scratch.p, t->gemm_scratch.n, st) || launch_softmax_ce(t->logits.p, pi_dev, n, t->losses.p, st) ||
        launch_value_loss_bwd(t->value.p, z_dev, t->h1pre.p, P + L.vfc2.off, n, C, t->slope, t->dpre.p, t->dh1pre.p, G + L.vfc2.off, t->losses.p, st) ||
        launch_sgemm(t->pfeat.p, t->logits.p, G + L.pfc.off, 128, 4672, n, 1, 0, t->gemm_scratch.p, t->gemm_scratch.n, st) ||            // dWp = pfeat^T dlogits
        launch_sgemm(t->logits.p, P + L.pfc.off, t->dpfeat.p, n, 128, 4672, 0, 1, t->gemm_scratch.p, t->gemm_scratch.n, st) ||           // dpfeat = dlogits Wp^T
        launch_sgemm(t->vfeat.p, t->dh1pre.p, G + L.vfc1.off, 64, C, n, 1, 0, t->gemm_scratch.p, t->gemm_scratch.n, st) ||               // dfc1 = vfeat^T dh1pre
        launch_sgemm(t->dh1pre.p, P + L.vfc1.off, t->dvfeat.p, n, 64, C, 0, 1, t->gemm_scratch.p, t->gemm_scratch.n, st) ||              // dvfeat = dh1pre fc1^T
        launch_head_bn_bwd_reduce(t->head_raw.p, t->vfeat.p, t->pfeat.p, t->dvfeat.p, t->dpfeat.p, n, t->hmr.p, t->slope, t->hsums.p, st) ||
        launch_head_conv_bwd(top, t->g[0].p, n, cj, C, t->head_raw.p, t->vfeat.p, t->pfeat.p, t->dvfeat.p, t->dpfeat.p, t->hmr.p, t->hsums.p,
                             P + L.vbn.off, P + L.pbn.off, P + L.vconv.off, P + L.pconv.off, t->slope, G + L.vconv.off, G + L.pconv.off, G + L.vbn.off,
                             G + L.pbn.off, st) ||
        tp_end(t, st))
        return -1;
    // ---------------------------------------------------------------- backward tower
    // Two branches.  The main stream carries the chain every layer waits for: BN backward (HBM-bound) -> data gradient (tower
    // kernel: it fills every SM's registers and shared memory).  The weight gradient of layer l hangs off that chain - it only
    // needs dY_l and the layer's input - so it runs on the auxiliary stream AFTER the layer's data gradient and thereby next to
    // the BN-backward passes of layer l - 1: its CTAs (1 per SM, 63 registers) leave room for the BN blocks, tensor pipe and HBM
    // are busy at the same time.  dY is double-buffered; the buffer of layer l is rewritten by layer l - 2 once wgrad(l) is done.
    // CB_TRAIN_SERIAL_WGRAD=1 (A/B runs) and per-launch profiling (clean kernel times) keep everything on one stream
    static const bool serial_env = getenv("CB_TRAIN_SERIAL_WGRAD") != nullptr;
    const bool serial_wgrad = serial_env || t->profile;
    cudaStream_t aux = serial_wgrad ? st : t->aux_stream;
    int cur = 0;
    for (int l = nl - 1; l >= 0; --l) {
        const int db = l & 1;
        const __nv_bfloat16* g = t->g[cur].p;
        double* sums = t->bsums.p + (size_t)l * 2 * N;
        const float* mr = t->mean_rstd.p + (size_t)l * 4 * N;
        __nv_bfloat16* dz = (l > 0 && (l & 1) == 0) ? t->dz.p : nullptr;  // second convolution of a block: its dz also flows down the skip
        const __nv_bfloat16* act = dz ? t->A[l].p : nullptr;              // without a residual add the sign of the pre-activation follows from y
        if (!serial_wgrad && l + 2 < nl) CB_CUDA_OK(cudaStreamWaitEvent(st, t->ev_wdone[l + 2], 0));   // wgrad(l + 2) has read this dY buffer
        if (tp_begin(t, TP_BN_BWD, st) || launch_bn_bwd_reduce(g, act, t->Y[l].p, n, cj, mr, t->slope, sums, st) ||
            launch_bn_bwd_apply(g, act, t->Y[l].p, t->dy[db].p, dz, n, cj, C, mr, P + L.bn[l].off, sums, t->slope, G + L.bn[l].off,
                                G + L.bn[l].off + C, st) ||
            tp_end(t, st))
            return -1;
        if (l > 0) {
            if (tp_begin(t, TP_DGRAD, st) ||
                conv_layer(t, nl + l, t->map_dy_conv[db], t->dy[db].p, t->g[cur ^ 1].p, (l & 1) ? t->dz.p : nullptr, nullptr, st) || tp_end(t, st))
                return -1;
            cur ^= 1;
        }
        if (!serial_wgrad) {
            CB_CUDA_OK(cudaEventRecord(t->ev_ready[l], st));
            CB_CUDA_OK(cudaStreamWaitEvent(aux, t->ev_ready[l], 0));
        }
        WgradParams wp{};
        wp.tm_a = l == 0 ? t->map_planes : t->map_A[l - 1];
        wp.tm_dy = t->map_dy_wg[db];
        wp.partial = t->wg_partial.p;
        wp.pairs = t->pairs;
        wp.cj_in = l == 0 ? 16 : cj;
        wp.n = N;
        wp.splits = l == 0 ? t->wg_splits_stem : t->wg_splits_res;
        if (tp_begin(t, TP_WGRAD, aux) || launch_wgrad(wp, aux) ||
            launch_wgrad_reduce(t->wg_partial.p, wp.splits, wp.cj_in * 8, N, l == 0 ? 119 : C, C, l == 0, G + L.w[l].off, aux) || tp_end(t, aux))
            return -1;
        if (!serial_wgrad) CB_CUDA_OK(cudaEventRecord(t->ev_wdone[l], aux));
    }
    if (!serial_wgrad) CB_CUDA_OK(cudaStreamWaitEvent(st, t->ev_wdone[0], 0));   // join: the weight gradients are serial on aux, layer 0 is the last
    return 0;
}

int cb_train_apply(cb_trainer* t, float lr, float momentum, int nesterov, float l2, float grad_scale, void* stream) {
    cb_device_guard cb_guard_(t->device);
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(trainer device) failed"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    t->last_l2 = l2;
    const float hp[5] = {lr, momentum, l2, grad_scale, nesterov ? 1.f : 0.f};
    CB_CUDA_OK(cudaMemcpyAsync(t->hp.p, hp, sizeof(hp), cudaMemcpyHostToDevice, st));  // pageable source: staged before the call returns
    return run_graphed(t, enqueue_apply, &t->graph_apply, &t->apply_calls, st);
}

static int enqueue_apply(cb_trainer* t, cudaStream_t st) {
    CB_CUDA_OK(cudaMemsetAsync(t->losses.p + 2, 0, 8, st));
    if (tp_begin(t, TP_OPT, st) ||
        launch_sgd(t->params, t->mom, t->grads, t->lay.n_params, t->lay.n_reg, t->hp.p, t->losses.p, st) || pack_weights(t, st) || tp_end(t, st))
        return -1;
    return 0;
}

// out[0] = mean squared error of the value head, out[1] = mean policy cross-entropy, out[2] = l2 * sum of squared kernels
// (as of the last cb_train_apply): the three terms Keras sums into `loss`
int cb_train_losses_sync(cb_trainer* t, double* out3, void* stream) {
    cb_device_guard cb_guard_(t->device);
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(trainer device) failed"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    double h[3];
    CB_CUDA_OK(cudaMemcpyAsync(h, t->losses.p, 24, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));
    out3[0] = h[0] / t->batch;
    out3[1] = h[1] / t->batch;
    out3[2] = h[2] * (double)t->last_l2;
    return 0;
}

// value [n] and policy logits gradient buffer are internal; the forward outputs of the last step for inspection / tests
int cb_train_outputs_sync(cb_trainer* t, float* value_host, void* stream) {
    cb_device_guard cb_guard_(t->device);
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(trainer device) failed"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA_OK(cudaMemcpyAsync(value_host, t->value.p, (size_t)t->batch * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int cb_train_profile(cb_trainer* t, int enable) {
    t->profile = enable != 0;
    t->ev_used = 0;
    return 0;
}
int cb_train_profile_read_sync(cb_trainer* t, double* total_ms, int* launches) {
    cb_device_guard cb_guard_(t->device);
    if (!cb_guard_.ok) { cb_set_error("cudaSetDevice(trainer device) failed"); return -1; }
    CB_CUDA_OK(cudaDeviceSynchronize());
    for (int k = 0; k < TP_STAGES; ++k) { total_ms[k] = 0; launches[k] = 0; }
    for (size_t i = 0; i + 1 < t->ev_used; i += 2) {
        float ms = 0;
        CB_CUDA_OK(cudaEventElapsedTime(&ms, t->ev[i], t->ev[i + 1]));
        total_ms[t->ev_stage[i / 2]] += ms;
        launches[t->ev_stage[i / 2]] += 1;
    }
    t->ev_used = 0;
    return 0;
}

}  // extern "C"
