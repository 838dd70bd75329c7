// Parameter blocks and launchers shared by the kernels (conv.cu, encode.cu, heads.cu, mcts.cu) and
// the C ABI (api.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "rules.cuh"

namespace cb {

struct ConvParams {
    const __nv_bfloat16* in;    // tile-native, cj_in chunks
    __nv_bfloat16* out;         // tile-native, n/8 chunks
    const __nv_bfloat16* skip;  // tile-native like out, or nullptr
    const __nv_bfloat16* wgt;   // [kc][half][tap][8][n/2][8], see conv_weight_index
    const float* scale;         // [n] gamma / sqrt(var + eps)
    const float* shift;         // [n] beta - mean * scale
    const float* head_w;        // [3][n] 1x1 head conv weights (value, policy0, policy1) or nullptr
    float* head_out;            // [boards][64][3]
    int pairs;                  // board pairs in the batch
    int kc;                     // K chunks of 64 (weight blocks)
    int kc_in;                  // distinct input chunks: A chunk of K chunk k is k % kc_in (stem: hi/lo weight halves re-read the planes)
    int cj_in;                  // input channels / 8
    int n;                      // output channels (multiple of 64, <= 256)
    float slope;                // LeakyReLU negative slope (Keras default 0.3)
};

// One layer of the tower kernel (device array, read by every role)
struct LayerDesc {
    const float* scale;         // [n] gamma / sqrt(var + eps)
    const float* shift;         // [n] beta - mean * scale
    int kc, kc_in, cj_in;       // as in ConvParams
    int in_buf, out_buf, skip_buf;  // indices into TowerParams::bufs / tm_act; skip_buf = -1: no residual add
    int w_map;                  // index into TowerParams::maps
    int pad_;
};

struct TowerParams {
    CUtensorMap tm_act[4];      // the activation buffers as rows of 128 B, box = one 64-channel chunk of a board pair (200 rows)
    const CUtensorMap* maps;    // device array: packed weights of each layer as rows of 128 B, box = conv_taps_per_box(n) tap blocks of n/2 rows
    const LayerDesc* layers;    // device array [nl]
    uint8_t* bufs[4];           // raw pointers of the same four buffers (epilogue stores, skip reads)
    const float* head_w;        // [3][n] 1x1 head conv weights fused into the LAST layer, or nullptr
    float* head_out;            // [boards][64][3]
    uint8_t* lo_buf;            // low half of the residual stream (bf16(x - bf16(x)), same tile-native layout as the activation buffers, only the
                                // epilogue reads / writes it, in place), or nullptr: stream rounded to bf16 at every block
    int fp8_probe;              // THROUGHPUT PROBE ONLY (env CB_TOWER_FP8_PROBE): e4m3 MMAs over half the k-chunks, results meaningless
    int lo_from_layer;          // the low half exists from the output of this layer on (0: whole tower); earlier blocks round the stream to bf16
    double* stats;              // training forward (one layer per launch, no skip): [2][n] per-channel sum / sum of squares of the
                                // bf16-rounded outputs, accumulated on top of what is there; nullptr otherwise
    int full_sectors;           // epilogue also stores the zero border cells beside the first / last column (whole 32-B DRAM sectors per board
                                // row): set when the activations do not stay in L2 anyway; costs store issue slots narrow layers cannot spare
    const int* n_boards_dev;    // search loop: number of boards actually in the batch this step (device-side, <= 2 * pairs); nullptr = 2 * pairs
    int* flags;                 // [2 * units] progress of every board pair: epoch * 256 + layers finished
    int epoch;                  // increases with every launch, so flags never need clearing
    int nl;                     // layers to run
    int pairs;                  // board pairs in the batch
    int n;                      // channels of every layer output (multiple of 64, <= 256)
    float slope;
};

// Packed weight layout [kc][half][tap][k/8][N/2][k%8]: per k-chunk of 64 input channels and half of the output
// channels nine K-major / no-swizzle tap blocks of N/2 rows x 128 B, contiguous, so that one TMA box can fetch one
// tap (wide layers) or three (narrow layers) as the B operand a CTA of the pair stages.
__host__ __device__ inline size_t conv_weight_index(int kc, int tap, int k /*0..63*/, int co, int n) {
    const int nh = n >> 1, half = co >= nh, c = co - half * nh;
    return ((size_t)(((kc * 2 + half) * 9 + tap) * 8 + (k >> 3)) * nh + c) * 8 + (k & 7);
}
// A weight stage (one ring slot, one "full" / "empty" barrier round trip, 12 MMAs) is always three taps of one k-chunk: the MMA-issuing warp
// pays its per-stage bookkeeping once per 1536 tensor cycles at N = 256 instead of once per 512 (ncu: with one-tap stages that warp was busy
// 92 % of the time and bounded the tensor pipe).  A TMA box holds at most 256 rows, so a stage is fetched as one box of three taps
// (N <= 128) or three boxes of one tap (N = 256).
__host__ __device__ inline int conv_taps_per_stage(int) { return 3; }
__host__ __device__ inline int conv_taps_per_box(int n) { return n <= 128 ? 3 : 1; }

struct HeadParams {
    const float* head_raw;   // [boards][64][3]
    const float* bn;         // [3][2] folded scale, shift for value conv, policy conv 0, policy conv 1
    const float* fc1;        // [64][C]
    const float* fc2;        // [C]
    int c;                   // filters
    float slope;
    float* value;            // [boards]
    float* pfeat;            // [boards][128]
    const int32_t* slot_list;  // search loop: board b of the (compacted) batch belongs to evaluation slot slot_list[b]: results are
    const int* n_boards_dev;   // scattered there; only *n_boards_dev boards are in the batch.  nullptr = identity / all boards
};

struct Node {  // 32 bytes; children of a node are contiguous, in legal-move order
    int32_t N;
    float W, Q, P;
    int32_t first_child;
    int32_t parent;
    int32_t pos_id;      // position of this node in the pos pool, -1 until first visited
    uint16_t move;       // move that leads here
    uint8_t n_children;
    int8_t term;         // 0 unknown / not terminal, 1 terminal value 0, 2 terminal value -1 (side to move is mated),
                         // 3 selected and waiting for its evaluation (only marked when leaves > 1)
};
static_assert(sizeof(Node) == 32, "Node must be 32 bytes");

constexpr int MCTS_MAX_PATH = 128;  // deeper paths fall back to walking parent links
enum { TREE_IDLE = 0, TREE_WAIT_ROOT = 1, TREE_WAIT_LEAF = 2, TREE_DONE = 3 };
enum { EVAL_NET = 0, EVAL_SYNTH = 1, EVAL_EXTERNAL = 2 };

constexpr int NODE_CHUNK = 4096;  // nodes a tree takes from the global pool at a time (one contended atomic per ~100 expansions)
constexpr int POS_CHUNK = 128;    // positions likewise
struct TreeState {  // 80 bytes
    int32_t root_node, root_pos;
    int32_t sims_done, sims_target;
    int32_t n_pending;   // leaves of this tree waiting for their evaluation (slots tree * leaves + 0 .. n_pending - 1)
    int32_t noise_off;   // first of this root's Gamma samples in SearchParams::noise (root children follow in legal-move order)
    int32_t rec_start;   // first position of this root's game record in the pool (the record is contiguous, one ply per slot)
    int32_t status;
    int32_t n_nodes, n_evals;
    int32_t noisy;       // add_exploration_noise applied at the root (P held in f64)
    int32_t max_len;     // config.max_game_length
    double cpuct;        // < 0: AlphaZero schedule log((N+19653)/19652)+1.25, else fixed (mcts.py:135-139)
    int32_t error;
    int32_t pad2_;
    int32_t node_next, node_end;  // this tree's current chunk of the node pool: [node_next, node_end) is free
    int32_t pos_next, pos_end;    // and of the position pool
};
static_assert(sizeof(TreeState) == 80, "TreeState must be 80 bytes");

// One in-flight leaf (evaluation slot).  Slot s = tree * leaves + i; the batch of a step is n_trees * leaves slots.
struct SlotState {
    int32_t leaf_node, leaf_pos, n_moves;
    int32_t path_len;    // nodes on the root..leaf path recorded in SearchParams::path (0 = not recorded: too deep)
};

struct SearchParams {
    Node* nodes; int32_t* node_count; int32_t max_nodes;
    Pos* pos; int32_t* pos_count; int32_t max_pos;
    TreeState* trees; int n_trees;
    SlotState* slots;          // [n_trees * leaves]
    int leaves;                // leaves per tree in flight: 1 = the reference's strictly sequential simulations (parity mode);
                               // > 1 = throughput mode with virtual loss (north_star), results differ from the sequential order
    uint16_t* pend_moves;      // [slots][MAX_MOVES]
    int32_t* frame_idx;        // [slots][8] newest first
    double* root_p64;          // [n_trees][MAX_MOVES]
    const double* noise;       // Gamma(0.3,1) samples drawn by the host (mcts.py:158), tree t's at [trees[t].noise_off ..]
    const double* cpuct_tab;   // [tab_len] log((N+19653)/19652)+1.25 from the host libm
    const double* sqrt_tab;    // [tab_len] pow(N, 0.5) from the host libm
    int tab_len;
    int eval_mode;
    const float* value;        // [slots]
    const float* pfeat;        // EVAL_NET: [slots][128]
    const float* wpt;          // EVAL_NET: policy Dense transposed [4672][128]
    const uint32_t* synth_hash;  // EVAL_SYNTH: [slots]
    const float* expvals;      // EVAL_EXTERNAL: [slots][4672] exp(logit) from the host (leaves == 1)
    const float* exp_table;    // [65536] f32 exp indexed by bf16 bits
    int32_t* counters;         // [0] evals, [1] sims, [2] errors
    int32_t* path;             // [slots][MCTS_MAX_PATH] node ids root..leaf of every pending leaf
    int32_t* active_list;      // [slots] the evaluation slots that hold a pending leaf after this step, in claim order: the next
    int32_t* active_count;     // step encodes / evaluates only these (a batch shrinks as trees finish their simulations)
    // EVAL_NET, fused step (tower + mcts_step only): the tree's warp finishes the heads of its own leaf in phase A and encodes the
    // leaf it selects at the end of phase B, so a step is two launches instead of four
    int32_t* slot_board;       // [slots] index of the slot's leaf in the evaluation batch it was claimed into (= its row in head_raw)
    const float* head_raw;     // [batch][64][3] raw 1x1 head convolutions of the tower (nullptr: heads come from sp.value / sp.pfeat)
    const float* head_bn;      // [3][2] folded scale / shift of the value conv and the two policy conv channels
    const float* fc1;          // [64][c] value Dense
    const float* fc2;          // [c]
    int head_c;                // filters
    float slope;               // LeakyReLU slope
    __nv_bfloat16* planes_act; // [batch] input planes of the NEXT step's batch (nullptr: a separate encode launch does it)
};

// Weight gradient of one 3x3 convolution (wgrad.cu)
struct WgradParams {
    CUtensorMap tm_a;    // layer input (tile-native, cj_in chunks per board pair) as rows of 128 B, box = 200 rows (64 channels of a pair)
    CUtensorMap tm_dy;   // gradient of the layer's raw output (n/8 chunks per pair), box = 25 * min(8, n/16) rows
    float* partial;      // [splits][9][cj_in*8][n] f32 partial sums
    int pairs;           // board pairs in the batch
    int cj_in;           // input channels / 8: 8, 16 or 32
    int n;               // output channels (padded): 64, 128 or 256
    int splits;          // split-K factor over the board pairs (wgrad_splits)
};

int launch_tower(const TowerParams& p, int num_sms, cudaStream_t stream);
int wgrad_splits(int num_sms, int cj_in);
int launch_wgrad(const WgradParams& p, cudaStream_t stream);
int launch_wgrad_reduce(const float* partial, int splits, int ci_pad, int n, int cin, int c, int stem, float* out, cudaStream_t stream);
int launch_wgrad_reference(const __nv_bfloat16* in, const __nv_bfloat16* dy, int pairs, int cj_in, int n, float* out, cudaStream_t stream);
int conv_make_row_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t box_rows);
int launch_conv3x3_reference(const ConvParams& p, cudaStream_t stream);
// training step (train.cu)
int launch_planes_from_images(const int16_t* img, int n, __nv_bfloat16* planes, cudaStream_t st);
int launch_bn_stats(const __nv_bfloat16* y, int pairs, int cj, double* stats, cudaStream_t st);
int launch_bn_apply(const __nv_bfloat16* y, const __nv_bfloat16* skip, __nv_bfloat16* out, int n, int cj, int c, const double* stats, const float* gamma,
                    const float* beta, float* mean_rstd, float* moving, float eps, float slope, float momentum, cudaStream_t st);
int launch_bn_bwd_reduce(const __nv_bfloat16* g, const __nv_bfloat16* a, const __nv_bfloat16* y, int n, int cj, const float* mean_rstd, float slope,
                         double* sums, cudaStream_t st);
int launch_bn_bwd_apply(const __nv_bfloat16* g, const __nv_bfloat16* a, const __nv_bfloat16* y, __nv_bfloat16* dy, __nv_bfloat16* dz, int n, int cj, int c,
                        const float* mean_rstd, const float* gamma, const double* sums, float slope, float* dgamma, float* dbeta, cudaStream_t st);
int launch_head_conv_fwd(const __nv_bfloat16* a, int n, int cj, const float* wv, const float* wp, int c, float* head_raw, cudaStream_t st);
int launch_head_stats(const float* head_raw, int n, double* stats, cudaStream_t st);
int launch_head_features_train(const float* head_raw, int n, const double* stats, const float* vbn, const float* pbn, float* vmoving, float* pmoving,
                               const float* fc1, const float* fc2, int c, float eps, float slope, float momentum, float* hmr, float* vfeat,
                               float* pfeat, float* h1pre, float* value, cudaStream_t st);
int launch_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int ta, int tb, float* scratch, size_t scratch_floats,
                 cudaStream_t st);
int launch_softmax_ce(float* logits, const float* pi, int n, double* losses, cudaStream_t st);
int launch_value_loss_bwd(const float* value, const float* z, const float* h1pre, const float* fc2, int n, int c, float slope, float* dpre, float* dh1pre,
                          float* dfc2, double* losses, cudaStream_t st);
int launch_head_bn_bwd_reduce(const float* head_raw, const float* vfeat, const float* pfeat, const float* dvfeat, const float* dpfeat, int n,
                              const float* hmr, float slope, double* hsums, cudaStream_t st);
int launch_head_conv_bwd(const __nv_bfloat16* a, __nv_bfloat16* gout, int n, int cj, int c, const float* head_raw, const float* vfeat, const float* pfeat,
                         const float* dvfeat, const float* dpfeat, const float* hmr, const double* hsums, const float* vbn, const float* pbn,
                         const float* wv, const float* wp, float slope, float* dwv, float* dwp, float* dvbn, float* dpbn, cudaStream_t st);
int launch_sgd(float* w, float* v, const float* g, size_t n, size_t n_reg, const float* hp_dev, double* losses, cudaStream_t st);
int launch_pack_conv(const float* w, int cin, int cin_pad, int c, int n, int stem, __nv_bfloat16* fwd, __nv_bfloat16* tr, cudaStream_t st);
int launch_fill(float* p, size_t n, float v, cudaStream_t st);

int launch_encode(const Pos* pos_base, const int32_t* frame_idx, int n, __nv_bfloat16* planes_act, int16_t* planes_i16, cudaStream_t stream,
                  const int32_t* slot_list = nullptr, const int* n_dev = nullptr);
int launch_head_features(const HeadParams& p, int boards, cudaStream_t stream);
int launch_policy_dense(const float* pfeat, const float* w, int boards, __nv_bfloat16* logits_bf16, float* logits_f32, cudaStream_t stream);
int launch_synth(const int16_t* planes_i16, int n, uint32_t seed, uint32_t* hash_out, float* value, __nv_bfloat16* logits_dense, cudaStream_t stream);
int launch_mcts_begin(const SearchParams& sp, cudaStream_t stream);
int launch_mcts_step(const SearchParams& sp, cudaStream_t stream);
int launch_leaf_priors(const Pos* pos, const int32_t* cur_idx, int n, const float* pfeat, const float* wpt, const float* exp_table,
                       uint16_t* out_moves, int16_t* out_actions, float* out_priors, int32_t* out_n, cudaStream_t stream);
int launch_pack_leaf_eval(int n, const float* value, const int32_t* cnt, const uint16_t* moves, const int16_t* actions, const float* priors,
                          uint8_t* out, cudaStream_t stream);
int launch_policy_softmax(const Pos* pos, const int32_t* cur_idx, int n, const __nv_bfloat16* logits, const float* exp_table,
                          uint16_t* out_moves, int16_t* out_actions, float* out_priors, int32_t* out_n, cudaStream_t stream);

}  // namespace cb
