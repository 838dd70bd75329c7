/* chessbot_b200 — C ABI of the B200-native self-play hot path.
 *
 * The reference (tomaz-piko/ChessBot) has no FFI layer: its boundary is the Python module surface of
 * main/mcts.py, model.py, gameimage.py, actionspace.py, game.py.  The Python modules
 * in chessbot_b200/ keep those names and signatures and bind exactly the entry points below through
 * ctypes; each declaration cites the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; device buffers are caller-owned (PyTorch allocates);
 * every device entry point takes the CUDA stream as a void* (cudaStream_t) and returns 0 on success,
 * <0 on error with the text available from cb_last_error().  Nothing synchronises the stream unless
 * its name ends in _sync.  Not thread-safe per context; one context per GPU.
 */
#ifndef CHESSBOT_B200_H
#define CHESSBOT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_NUM_ACTIONS 4672      /* config.py:16 */
#define CB_PLANES 119            /* config.py:21 with T=8 */
#define CB_PLANES_PAD 128        /* channel padding of the bf16 NHWC image */
#define CB_MAX_MOVES 224
#define CB_POS_BYTES 96          /* one position of a game record (csrc/rules.cuh: struct Pos) */

const char* cb_last_error(void);
const char* cb_version(void);

/* ---- host rules: the python-chess surface game.py / gameimage.py / puzzles.py call -------------- */
/* chess.Board(fen) / chess.Board()  (game.py:44; puzzles.py:50).  NULL fen = start position. */
typedef struct cb_board cb_board;
cb_board* cb_board_new(const char* fen);
void cb_board_free(cb_board* b);
cb_board* cb_board_copy(const cb_board* b);                 /* Board.copy(), keeps the stack (game.py:137) */
int cb_board_push_uci(cb_board* b, const char* uci);        /* Board.push_uci (game.py:101); -1 if illegal */
int cb_board_pop(cb_board* b);                              /* Board.pop (gameimage.py:178); -1 if empty */
int cb_board_ply(const cb_board* b);                        /* len(board.move_stack) (game.py:25) */
int cb_board_turn(const cb_board* b);                       /* board.turn: 1 white, 0 black (game.py:145) */
int cb_board_halfmove_clock(const cb_board* b);             /* gameimage.py:146 */
int cb_board_castling(const cb_board* b);                   /* cleaned rights: bit0 WK, 1 WQ, 2 BK, 3 BQ (gameimage.py:121-127) */
int cb_board_ep_square(const cb_board* b);                  /* raw ep square or -1 */
uint64_t cb_board_pieces_mask(const cb_board* b, int piece_type /*1..6*/, int color); /* board.pieces (gameimage.py:51,54) */
int cb_board_legal_moves(const cb_board* b, uint16_t* moves /*[CB_MAX_MOVES]*/);      /* ordered like python-chess (game.py:85) */
int cb_board_move_stack(const cb_board* b, uint16_t* moves, int cap);                 /* game.py:16 */
int cb_board_outcome(const cb_board* b, int claim_draw);    /* board.outcome(claim_draw=) (game.py:57-58): cb::Outcome code */
int cb_board_is_repetition(const cb_board* b, int count);   /* gameimage.py:73 */
int cb_board_is_check(const cb_board* b);
int cb_board_fen(const cb_board* b, char* out, int cap);
int cb_board_mirror(cb_board* b);                           /* Board.apply_mirror (gameimage.py:39) */
uint64_t cb_board_perft(const cb_board* b, int depth);
/* Game record for the device: the last `n` positions (oldest first), enough for T=8 frames and for
 * repetition counting back to the last irreversible move.  Returns n, or the required n if cap is too small. */
int cb_board_export_record(const cb_board* b, void* pos_out /*[cap][CB_POS_BYTES]*/, int cap);
/* the same for n boards at once (roots of a batched search): records concatenated, rec_len[n] positions each, n_legal[n] =
 * legal moves of each board (length of its root noise, mcts.py:158; may be NULL).  Returns the total number of positions,
 * or the total required if cap is too small. */
int cb_boards_export_records(const cb_board* const* boards, int n, void* pos_out /*[cap][CB_POS_BYTES]*/, int cap,
                             int32_t* rec_len, int32_t* n_legal);
/* actionspace.uci_to_action (actionspace.py:116-180): index 0..4671, -1 where the reference returns None */
int cb_uci_to_action(const char* uci, int player_white);
int cb_move_to_action(uint16_t move, int player_white);
int cb_move_to_uci(uint16_t move, char out[6]);

/* ---- device context --------------------------------------------------------------------------- */
/* Device memory.  The tensors a call takes or returns are always caller-owned device pointers.  What the library keeps between calls
 * (packed weights and activations of cb_net_*, the node / position pools of cb_search_*, the trainer's buffers) it takes from the
 * allocator installed here - the Python host installs PyTorch's caching allocator, so every byte of device memory is owned by
 * PyTorch - and from cudaMalloc only if none is installed.  Install before the first context is created; process-wide.  Passing NULLs
 * afterwards detaches the host allocator (interpreter shutdown): blocks still held are then left to process exit. */
int cb_set_device_allocator(void* (*alloc)(size_t bytes, int device, void* user), void (*release)(void* ptr, void* user), void* user);
/* what cb_search_create(max_trees, max_nodes, max_positions, max_sims) followed by cb_search_set_leaves(leaves) will ask for */
size_t cb_search_workspace_bytes(int max_trees, int max_nodes, int max_positions, int max_sims, int leaves);
typedef struct cb_ctx cb_ctx;
cb_ctx* cb_ctx_create(int device);                 /* one per GPU / process; NULL + cb_last_error() on failure */
void cb_ctx_destroy(cb_ctx* ctx);
int cb_ctx_sm_count(const cb_ctx* ctx);
int cb_ctx_device(const cb_ctx* ctx);             /* the CUDA device of this context; every entry point makes it current */
/* exp() used for priors (mcts.py:94): f32 table indexed by the bf16 bit pattern of the logit.  The Python
 * host fills it with np.exp so priors are bit-identical to NumPy on that host; default = expf. */
int cb_set_exp_table(cb_ctx* ctx, const float* table_host /*[65536]*/);

/* ---- gameimage.board_to_image, batched (gameimage.py:149-192) ----------------------------------- */
/* pos_dev: game-record positions; frame_idx_dev[n][8]: indices of the frames, newest first, -1 = absent.
 * planes_act_dev: bf16 planes in the convolution's tile-native layout (cb_planes_act_bytes(n) bytes, must
 * have been zeroed once: the border cells are never written); planes_i16_dev: the reference's own
 * int16 [n][8][8][119] image.  Either output may be NULL. */
size_t cb_planes_act_bytes(int n);
int cb_encode(cb_ctx* ctx, const void* pos_dev, const int32_t* frame_idx_dev, int n, void* planes_act_dev,
              int16_t* planes_i16_dev, void* stream);

/* ---- model.py: generate_model / predict_fn forward graph (model.py:19-56,73-80) ------------------- */
int cb_net_create(cb_ctx* ctx, int filters, int blocks, int max_batch, float leaky_slope, float bn_eps);
/* Keras-layout float32 tensors by name: "stem.w" [3,3,119,C] HWIO, "stem.bn" [4,C] (gamma,beta,mean,var),
 * "res{i}.w1|w2" [3,3,C,C], "res{i}.bn1|bn2" [4,C], "value.conv" [C,1], "value.bn" [4,1], "value.fc1" [64,C],
 * "value.fc2" [C,1], "policy.conv" [C,2], "policy.bn" [4,2], "policy.fc" [128,4672]. */
int cb_net_set_tensor(cb_ctx* ctx, const char* name, const float* host, size_t count);
int cb_net_finalize(cb_ctx* ctx);                  /* fold BN, pack bf16 weights, upload */
/* How the residual stream x_{i+1} = LReLU(x_i + branch_i) of model.py:19-28 is kept between blocks.  1 (default): bf16 hi + bf16 lo
 * (~16 mantissa bits: at 20x256 the 2e-2 tolerance on the value holds with margin, profiles/r2_precision_variants.json);
 * 0: rounded to bf16 at every block (round-1 behaviour, one activation buffer less).  The reference's TF-TRT FP32 graph has no
 * counterpart (it never rounds).  k >= 2: hi + lo only from the input of residual block k - 1 on (earlier blocks round: a measured
 * precision / speed trade, profiles/r2_precision_variants.md).  Call between cb_net_create and cb_net_finalize. */
int cb_net_set_stream_precision(cb_ctx* ctx, int hi_lo);
/* planes (tile-native bf16) -> value f32[n] and policy logits [n][4672] (bf16 and/or f32, either may be NULL) */
int cb_net_forward(cb_ctx* ctx, const void* planes_act_dev, int n, float* value_dev, void* logits_bf16_dev,
                   float* logits_f32_dev, void* stream);
/* the reference's mock-network contract (tests/mcts_v2.py:51-67) as a device kernel: hash of the int16 planes */
int cb_synth_forward(cb_ctx* ctx, const int16_t* planes_i16_dev, int n, uint32_t seed, float* value_dev,
                     void* logits_bf16_dev, void* stream);

/* ---- mcts.expand's gather + exp + normalise, batched (mcts.py:91-101; actionspace.py:116-180) ------- */
/* per position i: ordered legal moves of pos_dev[cur_idx_dev[i]], their action indices and priors.
 * outputs are [n][CB_MAX_MOVES] (moves uint16: from | to<<6 | promo<<12) and out_n[n]. */
int cb_policy_softmax(cb_ctx* ctx, const void* pos_dev, const int32_t* cur_idx_dev, int n, const void* logits_bf16_dev,
                      uint16_t* out_moves_dev, int16_t* out_actions_dev, float* out_priors_dev, int32_t* out_n_dev,
                      void* stream);

/* ---- batched leaf evaluation outside a tree (BASELINE config 2; mcts.expand, mcts.py:80-103, for n positions at once) ---------- */
/* encode + tower + heads + legal-move gather / exp / normalise WITHOUT the dense [n][4672] logits: the policy Dense is evaluated for the
 * legal moves only (same arithmetic, same bits as cb_net_forward + cb_policy_softmax).  cur_idx_dev[n]: index of each position in
 * pos_dev (= frame_idx_dev[i][0]); planes_act_dev: zeroed scratch of cb_planes_act_bytes(n); outputs [n][CB_MAX_MOVES] rows. */
int cb_leaf_eval(cb_ctx* ctx, const void* pos_dev, const int32_t* frame_idx_dev, const int32_t* cur_idx_dev, int n, void* planes_act_dev,
                 float* value_dev, uint16_t* moves_dev, int16_t* actions_dev, float* priors_dev, int32_t* n_moves_dev, void* stream);
/* the same end to end from HOST game records (concatenated cb_board_export_record blocks, rec_len[n] positions each): pinned staging,
 * one H2D, the kernels, a packed result, two small D2H.  out_host: value f32[n] | n_moves i32[n] | offsets i32[n+1] | priors f32[T] |
 * moves u16[T] | actions i16[T]; returns T (<0: error).  out_cap_bytes >= cb_leaf_eval_packed_bytes(n, T); n * CB_MAX_MOVES always fits. */
size_t cb_leaf_eval_packed_bytes(int n, int total_moves);
int cb_leaf_eval_host_sync(cb_ctx* ctx, const void* records_host, const int32_t* rec_len, int n, void* out_host, size_t out_cap_bytes, void* stream);

/* ---- mcts.run_mcts over many concurrent trees (mcts.py:38-161) ---------------------------------- */
#define CB_EVAL_NET 0        /* the network loaded with cb_net_* */
#define CB_EVAL_SYNTH 1      /* synthetic hash network */
#define CB_EVAL_EXTERNAL 2   /* host callable: cb_search_pending_planes / cb_search_step_external */
int cb_search_create(cb_ctx* ctx, int max_trees, int max_nodes, int max_positions, int max_sims);
/* leaves of one tree evaluated per step.  1 (default) = the reference's strictly sequential simulations: visit counts are
 * the reference's.  > 1 = throughput mode for few games (north_star: "virtual-loss ... over many concurrent trees"): a tree
 * selects up to `leaves` leaves per step, each lowering its path by one lost game until its evaluation arrives; the
 * evaluation batch is n_trees * leaves.  Call between cb_search_create and cb_search_begin. */
int cb_search_set_leaves(cb_ctx* ctx, int leaves);
/* roots: concatenated game records (cb_board_export_record), rec_len[n_trees] positions each (host memory).
 * sims[n], cpuct[n] (<0 = AlphaZero schedule, mcts.py:135-139), noise: NULL or [n][CB_MAX_MOVES] Gamma(0.3,1)
 * samples per tree (mcts.py:158), has_noise[n] flags; max_game_length: config.py:33. */
int cb_search_begin(cb_ctx* ctx, int n_trees, const void* records_host, const int32_t* rec_len, const int32_t* sims,
                    const double* cpuct, const uint8_t* has_noise, const double* noise, int max_game_length, int eval_mode,
                    uint32_t synth_seed, void* stream);
/* the same with the root noise in CSR form: tree t's samples are noise_flat[noise_off[t] .. noise_off[t+1]) (an empty range = no noise
 * for that tree; noise_flat NULL = none at all).  Neither variant synchronises with the device: inputs are copied to pinned staging
 * owned by the context before the call returns. */
int cb_search_begin_csr(cb_ctx* ctx, int n_trees, const void* records_host, const int32_t* rec_len, const int32_t* sims,
                        const double* cpuct, const double* noise_flat, const int64_t* noise_off, int max_game_length, int eval_mode,
                        uint32_t synth_seed, void* stream);
/* one batch step: encode pending leaves, evaluate them, expand + backup, select the next leaves */
int cb_search_step(cb_ctx* ctx, void* stream);
int cb_search_run(cb_ctx* ctx, int steps, void* stream);     /* `steps` x cb_search_step, no host sync */
/* external evaluation: int16 planes of the pending leaf of every tree -> host; then values + exp(logits) back */
int cb_search_pending_planes_sync(cb_ctx* ctx, int16_t* planes_host /*[n_trees][8][8][119]*/, uint8_t* waiting_host /*[n_trees]*/, void* stream);
int cb_search_step_external(cb_ctx* ctx, const float* value_host /*[n_trees]*/, const float* expvals_host /*[n_trees][4672]*/, void* stream);
/* counters[4]: trees not finished, evaluations, simulations, errors (node/position pool exhausted) */
int cb_search_status_sync(cb_ctx* ctx, int64_t* counters, void* stream);
/* root statistics of every tree: [n_trees][CB_MAX_MOVES] arrays + n_children[n_trees], root_n[n_trees], nodes[n_trees] (host) */
int cb_search_read_roots_sync(cb_ctx* ctx, uint16_t* moves, int32_t* n, float* w, float* q, double* p, int32_t* n_children,
                              int32_t* root_n, int32_t* tree_nodes, void* stream);
/* the same, packed: the children of tree t at [child_off[t], child_off[t+1]) of every section (child_off[n_trees + 1], host; the caller
 * knows every root's legal-move count from cb_boards_export_records), one device-to-host copy into pinned staging.  out_host holds
 * cb_search_roots_compact_bytes(n_trees, T = child_off[n_trees]) bytes:
 * P f64[T] | N i32[T] | W f32[T] | Q f32[T] | n_children i32[n] | root_N i32[n] | tree_nodes i32[n] | state i32[n] | moves u16[T]
 * (state = tree status | error << 8; train.py:13-19 reads child.N of the root; mcts.select_move, mcts.py:142-153, the visit counts).
 * _async queues the read-back (+ an event) behind the search steps already on the stream and returns at once; _wait blocks on that
 * event only and also delivers the counters of cb_search_status_sync, so a host can collect one search while the next one (another
 * context on the same stream) is already running - no device-wide or stream-wide synchronisation. */
size_t cb_search_roots_compact_bytes(int n_trees, int total_children);
int cb_search_read_roots_compact_sync(cb_ctx* ctx, const int32_t* child_off, void* out_host, void* stream);
int cb_search_read_roots_compact_async(cb_ctx* ctx, const int32_t* child_off, void* stream);
int cb_search_read_roots_compact_wait(cb_ctx* ctx, void* out_host, int64_t* counters /*[4] or NULL*/);

/* ---- train.py training step (train.py:150-164 -> model.fit; loss / optimizer of model.py:59-70) ------------------------ */
/* One SGD step on a batch of replay samples (train.py:104-110: int16 images, terminal values z, visit distributions pi):
 * loss = mean_squared_error(value, z) + CategoricalCrossentropy(from_logits)(logits, pi) + l2 * sum(kernel^2);
 * BatchNormalization in training mode (batch statistics, moving averages with `bn_momentum`); SGD with (Nesterov) momentum.
 * Parameters, gradients and momentum are three caller-owned flat f32 device buffers of cb_train_param_count() elements
 * (the first *n_regularized are the conv / dense kernels the l2 term applies to, then the BN gamma / beta).  With several
 * GPUs the caller all-reduces the gradient buffer (NCCL) between cb_train_forward_backward and cb_train_apply and
 * passes grad_scale = 1 / world_size.  Tensor names and layouts are those of cb_net_set_tensor. */
typedef struct cb_trainer cb_trainer;
size_t cb_train_param_count(int filters, int blocks, size_t* n_regularized);
cb_trainer* cb_train_create(cb_ctx* ctx, int filters, int blocks, int batch, float leaky_slope, float bn_eps, float bn_momentum,
                            float* params_dev, float* grads_dev, float* momentum_dev);
void cb_train_destroy(cb_trainer* t);
/* what: 0 = parameter (BN: gamma, beta, moving mean, moving variance), 1 = gradient (get only), 2 = momentum; both synchronise */
int cb_train_set_tensor(cb_trainer* t, const char* name, int what, const float* host, size_t count, void* stream);
int cb_train_get_tensor(cb_trainer* t, const char* name, int what, float* host, size_t count, void* stream);
int cb_train_forward_backward(cb_trainer* t, const int16_t* images_dev /*[n][8][8][119]*/, const float* z_dev /*[n]*/,
                              const float* pi_dev /*[n][4672]*/, int n, void* stream);
int cb_train_apply(cb_trainer* t, float lr, float momentum, int nesterov, float l2, float grad_scale, void* stream);
int cb_train_losses_sync(cb_trainer* t, double* out3 /* value mse, policy ce, l2 term */, void* stream);
int cb_train_outputs_sync(cb_trainer* t, float* value_host /*[n]*/, void* stream);
#define CB_TRAIN_PROF_STAGES 8   /* conv forward, BN forward, heads + losses, BN backward, wgrad, dgrad, optimizer + packing, input */
int cb_train_profile(cb_trainer* t, int enable);
int cb_train_profile_read_sync(cb_trainer* t, double* total_ms, int* launches);

/* ---- measurement: CUDA events (on the launching stream) around every kernel launch of the entry points above ---- */
#define CB_PROF_ENCODE 0
#define CB_PROF_CONV 1
#define CB_PROF_HEADS 2
#define CB_PROF_MCTS 3
#define CB_PROF_POLICY_DENSE 4
#define CB_PROF_POLICY_SOFTMAX 5
#define CB_PROF_STAGES 6
int cb_profile(cb_ctx* ctx, int enable);
int cb_profile_read_sync(cb_ctx* ctx, double* total_ms /*[CB_PROF_STAGES]*/, int* launches /*[CB_PROF_STAGES]*/);   /* sums and resets */

/* ---- test hooks (not on the product path) ------------------------------------------------------- */
/* one 3x3 conv + BN + (skip) + LeakyReLU layer of the loaded net on tile-native buffers: layer 0 = stem,
 * 2i+1 / 2i+2 = first / second conv of block i.  use_reference = plain CUDA-core kernel on the same inputs. */
int cb_debug_conv_layer(cb_ctx* ctx, int layer, const void* in_act_dev, const void* skip_act_dev, void* out_act_dev,
                        int n_boards, int use_reference, void* stream);

/* weight gradient of one 3x3 convolution (csrc/wgrad.cu) on tile-native buffers: dW[tap][ci][co] = sum over boards and
 * squares of in[.., ci] shifted by the tap times dy[.., co]; out_dev f32 [9][cin_pad][n], padded channels included. */
int cb_debug_wgrad(cb_ctx* ctx, const void* in_act_dev, const void* dy_act_dev, int n_boards, int cin_pad, int n,
                   float* out_dev, int use_reference, void* stream);

#ifdef __cplusplus
}
#endif
#endif
