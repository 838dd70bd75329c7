// K6/K7/K8 — PUCT select, device chess rules at the leaf, expand and backup (mcts.py:38-161) for
// many concurrent trees over a GPU-resident node pool.  One warp per tree.  With one in-flight leaf per
// tree (leaves = 1, the default) every tree runs the reference's strictly sequential simulation order, the
// batching is ACROSS trees (batch = number of games), so visit counts are those of the reference.  With
// leaves > 1 (throughput mode for few games) a tree selects several leaves per step under virtual loss.
//
// One launch of mcts_step_kernel per batch step does, per tree:
//   A. if a leaf evaluation is pending: priors for its legal moves (K5: gather, exp table, NumPy
//      pairwise sum, divide), children appended to the node pool, value backed up to the root;
//   B. then simulations until the next leaf that needs the network: descend by arg-max ucb
//      (f32 arithmetic without FMA contraction; f64 at a noisy root; ties -> larger UCI string),
//      make the move (rules.cuh), test outcome(claim_draw=True); terminal leaves are backed up
//      immediately (the reference re-evaluates them on every visit, mcts.py:67-74) and the descent
//      restarts; a non-terminal leaf records its ordered legal moves + the 8 frames to encode.
#include "kernels.cuh"
#include "encode.cuh"
#include "policy.cuh"
#include "rules_warp.cuh"

namespace cb {





__device__ __forceinline__ const Pos* prev_pos(const Pos* base, const Pos* q) {
    const int32_t pi = (int32_t)q->pad_[0];
    return pi >= 0 ? base + pi : nullptr;
}
struct PoolPrev {
    const Pos* base;
    __device__ const Pos* operator()(const Pos* q) const { return prev_pos(base, q); }
};

__device__ __forceinline__ uint32_t fmix32_d(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

// frames for the encoder: walk the record newest first, stop after a position with an empty move stack
__device__ void fill_frames(const SearchParams& sp, int slot, int pos_id) {
    int32_t* fi = sp.frame_idx + (size_t)slot * 8;
    int cur = pos_id;
    for (int t = 0; t < 8; ++t) {
        fi[t] = cur;
        if (cur >= 0) {
            const Pos& q = sp.pos[cur];
            cur = q.ply == 0 ? -1 : (int32_t)q.pad_[0];
        }
    }
}

// mcts.update (mcts.py:112-117), f32 accumulate, sign flip per ply
__device__ void backup(Node* nodes, int node, float v) {
    while (node >= 0) {
        Node& nd = nodes[node];
        nd.N += 1;
        nd.W = __fadd_rn(nd.W, v);
        nd.Q = __fdiv_rn(nd.W, (float)nd.N);
        v = -v;
        node = nd.parent;
    }
}

// the same update for a recorded root..leaf path: one node per lane (every node of a path is distinct, each
// receives exactly one f32 add, so the result does not depend on the order); `v` is the value for the LEAF
__device__ __forceinline__ void backup_path(Node* nodes, const int32_t* path, int len, float v, int lane) {
    for (int d = lane; d < len; d += 32) {
        Node& nd = nodes[path[d]];
        const int n = nd.N + 1;
        const float w = __fadd_rn(nd.W, ((len - 1 - d) & 1) ? -v : v);
        nd.N = n;
        nd.W = w;
        nd.Q = __fdiv_rn(w, (float)n);
    }
}

// Bump allocation inside a per-tree chunk; only a new chunk touches the global counter (every warp hammering one
// address costs ~30 cycles per warp at the L2 atomic unit: 1024 trees -> tens of microseconds per step).
// Returns the first index of `n` contiguous free slots, or -1 when the pool is exhausted.
__device__ __forceinline__ int chunk_alloc(int32_t& next, int32_t& end, int n, int chunk, int32_t* global_count, int max_total) {
    if (next + n > end) {
        const int take = n > chunk ? n : chunk;
        const int base = atomicAdd(global_count, take);
        if (base + take > max_total) return -1;
        next = base;
        end = base + take;
    }
    const int r = next;
    next += n;
    return r;
}

// Throughput mode (leaves > 1): a selected leaf lowers every node of its path by one lost game (N += 1, W -= 1) until its
// evaluation arrives, so that the next selections of the same step go elsewhere ...
__device__ __forceinline__ void virtual_loss_path(Node* nodes, const int32_t* path, int len, int lane) {
    for (int d = lane; d < len; d += 32) {
        Node& nd = nodes[path[d]];
        const int n = nd.N + 1;
        const float w = __fadd_rn(nd.W, -1.0f);
        nd.N = n;
        nd.W = w;
        nd.Q = __fdiv_rn(w, (float)n);
    }
}
// ... and the evaluation replaces the lost game by the real value (the visit is already counted)
__device__ __forceinline__ void settle_path(Node* nodes, const int32_t* path, int len, float v, int lane) {
    for (int d = lane; d < len; d += 32) {
        Node& nd = nodes[path[d]];
        const float w = __fadd_rn(nd.W, __fadd_rn(1.0f, ((len - 1 - d) & 1) ? -v : v));
        nd.W = w;
        nd.Q = __fdiv_rn(w, (float)nd.N);
    }
}

// K3 as a warp routine: the heads of ONE board (model.py:43-56 after the fused 1x1 convolutions), bit for bit what head_features_kernel
// (heads.cu, one block of 256 threads per board) computes - same fmaf chains, the same 32-lane xor butterflies per group of 32
// channels, the same sequential sum of the eight group totals - so a position's value and policy features do not depend on which of
// the two produced them.  pfeat (128) and v0 (64) live in the warp's shared-memory scratch; every lane returns the value.
__device__ __forceinline__ float head_features_warp(const float* __restrict__ raw /*[64][3]*/, const float* __restrict__ bn, const float* __restrict__ fc1,
                                                    const float* __restrict__ fc2, int c, float slope, float* s_v0 /*[64]*/, float* s_pfeat /*[128]*/,
                                                    int lane) {
    auto lrelu = [slope](float x) { return x > 0.f ? x : x * slope; };
    for (int t = lane; t < 64; t += 32) s_v0[t] = lrelu(fmaf(raw[t * 3], bn[0], bn[1]));
    for (int t = lane; t < 128; t += 32) {
        const int sq = t >> 1, ch = t & 1;
        s_pfeat[t] = lrelu(fmaf(raw[sq * 3 + 1 + ch], bn[2 + 2 * ch], bn[3 + 2 * ch]));
    }
    __syncwarp();
    float s = 0.f;
    for (int g = 0; g < 8; ++g) {  // group g = warp g of the block kernel: channels 32 g .. 32 g + 31
        const int t = g * 32 + lane;
        float part = 0.f;
        if (t < c) {
            float h = 0.f;
            for (int sq = 0; sq < 64; ++sq) h = fmaf(s_v0[sq], fc1[sq * c + t], h);
            part = lrelu(h) * fc2[t];
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        s += __shfl_sync(0xffffffffu, part, 0);  // the block kernel keeps lane 0's total of every group
    }
    return tanhf(s);
}

constexpr int MCTS_WARPS = 4;
constexpr int MCTS_TERMINAL_SIMS_PER_STEP = 4;

__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_step_kernel(SearchParams sp) {
    __shared__ float s_p[MCTS_WARPS][MAX_MOVES];
    __shared__ uint16_t s_m[MCTS_WARPS][MAX_MOVES];
    __shared__ int32_t s_path[MCTS_WARPS][MCTS_MAX_PATH];
    __shared__ u64 s_hh[MCTS_WARPS][MCTS_HIST];      // key hashes / position ids of the reversible history of the leaf being set up
    __shared__ int32_t s_hp[MCTS_WARPS][MCTS_HIST];
    __shared__ float s_pf[MCTS_WARPS][128];          // fused step: policy features / value-head inputs of the leaf being expanded
    __shared__ float s_v0[MCTS_WARPS][64];
    __shared__ u64 s_mask[MCTS_WARPS][112];          // fused step: encoder scratch for the leaf selected at the end
    __shared__ int s_scal[MCTS_WARPS][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tree = blockIdx.x * MCTS_WARPS + warp;
    if (tree >= sp.n_trees) return;
    TreeState& ts = sp.trees[tree];
    const int status = ts.status;
    if (status == TREE_IDLE || status == TREE_DONE || ts.error) return;
    float* pv = s_p[warp];
    const int L = sp.leaves;
    const bool vl = L > 1;  // throughput mode: virtual loss, several leaves of this tree per step

    // ------------------------------------------------------------------ A. expand + backup every evaluated leaf
    const int n_pending = ts.n_pending;
    for (int li = 0; li < n_pending; ++li) {
        const int slot = tree * L + li;
        const SlotState sl = sp.slots[slot];
        const int n = sl.n_moves, leaf = sl.leaf_node;
        const uint16_t* moves = sp.pend_moves + (size_t)slot * MAX_MOVES;
        const int turn = sp.pos[sl.leaf_pos].turn;
        const bool fused = sp.eval_mode == EVAL_NET && sp.head_raw != nullptr;
        float value_fused = 0.f;
        if (fused) {
            value_fused = head_features_warp(sp.head_raw + (size_t)sp.slot_board[slot] * 192, sp.head_bn, sp.fc1, sp.fc2, sp.head_c, sp.slope,
                                             s_v0[warp], s_pf[warp], lane);
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) {
            const int a = move_to_action(moves[i], turn);
            float e;
            if (sp.eval_mode == EVAL_NET) {
                const float4* wr = reinterpret_cast<const float4*>(sp.wpt + (size_t)a * 128);
                const float4* pf = fused ? reinterpret_cast<const float4*>(s_pf[warp]) : reinterpret_cast<const float4*>(sp.pfeat + (size_t)slot * 128);
                float acc = 0.f;
#pragma unroll 8
                for (int k = 0; k < 32; ++k) {
                    const float4 w = wr[k], f = pf[k];
                    acc = fmaf(f.x, w.x, acc); acc = fmaf(f.y, w.y, acc); acc = fmaf(f.z, w.z, acc); acc = fmaf(f.w, w.w, acc);
                }
                e = sp.exp_table[__bfloat16_as_ushort(__float2bfloat16_rn(acc))];
            } else if (sp.eval_mode == EVAL_SYNTH) {
                const uint32_t u = fmix32_d(sp.synth_hash[slot] ^ (((uint32_t)a + 1u) * 0x9E3779B1u));
                const float logit = (float)((int)((u >> 16) & 0xFF) - 128) / 32.0f;
                e = sp.exp_table[__bfloat16_as_ushort(__float2bfloat16_rn(logit))];
            } else {
                e = sp.expvals[(size_t)slot * 4672 + a];
            }
            pv[i] = e;
        }
        __syncwarp();
        const float sum = numpy_sum_f32_warp(pv, n, lane);
        int base = 0;
        if (lane == 0) {
            base = chunk_alloc(ts.node_next, ts.node_end, n, NODE_CHUNK, sp.node_count, sp.max_nodes);
            if (base < 0) { ts.error = 1; atomicAdd(&sp.counters[2], 1); }
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base < 0) return;
        const bool is_root = status == TREE_WAIT_ROOT;
        for (int i = lane; i < n; i += 32) {
            Node c;
            c.N = 0; c.W = 0.f; c.Q = 0.f;
            c.P = __fdiv_rn(pv[i], sum);
            c.first_child = -1; c.parent = leaf; c.pos_id = -1;
            c.move = moves[i]; c.n_children = 0; c.term = 0;
            sp.nodes[base + i] = c;
            if (is_root && ts.noisy) {
                // mcts.py:161: P*(1-frac) in f32, noise*frac in f64, sum in f64
                const float p75 = __fmul_rn(c.P, 0.75f);
                sp.root_p64[(size_t)tree * MAX_MOVES + i] =
                    __dadd_rn((double)p75, __dmul_rn(sp.noise[(size_t)ts.noise_off + i], 0.25));
            }
        }
        __syncwarp();
        const int plen = sl.path_len;
        const float v = -(fused ? value_fused : sp.value[slot]);  // update(node, -value), mcts.py:74
        if (lane == 0) {
            Node& lf = sp.nodes[leaf];
            lf.first_child = base;
            lf.n_children = (uint8_t)n;
            lf.term = 0;
            ts.n_nodes += n;
            ts.n_evals += 1;
            atomicAdd(&sp.counters[0], 1);
            if (!is_root) {
                if (plen == 0) backup(sp.nodes, leaf, v);  // path too deep to record: walk the parent links (never with virtual loss)
                ts.sims_done += 1;
                atomicAdd(&sp.counters[1], 1);
            }
            __threadfence_block();
        }
        __syncwarp();
        if (!is_root && plen > 0) {
            const int32_t* path = sp.path + (size_t)slot * MCTS_MAX_PATH;
            if (vl) settle_path(sp.nodes, path, plen, v, lane);
            else backup_path(sp.nodes, path, plen, v, lane);
        }
        __syncwarp();
    }

    // ------------------------------------------------------------------ B. simulate until `leaves` leaves wait for the network
    int sims_done = __shfl_sync(0xffffffffu, ts.sims_done, 0);
    const int sims_target = ts.sims_target;
    const bool noisy = ts.noisy != 0;
    const double cpuct_fixed = ts.cpuct;
    int n_new = 0, n_term = 0;
    while (sims_done + n_new < sims_target && n_new < L) {
        int node = ts.root_node;
        int parent_pos = ts.root_pos;
        int32_t* path = s_path[warp];
        int depth = 0;  // nodes recorded in path[] (root first); -1: too deep, not recorded
        // ---- select (mcts.py:62-64,119-140).  The header of the node being expanded (N, first_child, n_children, ...)
        // travels in registers: it is the child entry the previous level has just read, so a level costs ONE
        // dependent memory access (its children), not two.
        Node nd = sp.nodes[node];
        while (true) {
            if (lane == 0 && depth >= 0) path[depth] = node;
            depth = (depth >= 0 && depth + 1 < MCTS_MAX_PATH) ? depth + 1 : -1;
            if (nd.n_children == 0) break;
            if (nd.pos_id >= 0) parent_pos = nd.pos_id;
            const int np = nd.N < sp.tab_len ? nd.N : sp.tab_len - 1;
            const double c = cpuct_fixed < 0 ? sp.cpuct_tab[np] : cpuct_fixed;
            const double sq = sp.sqrt_tab[np];
            const bool f64 = noisy && node == ts.root_node;
            // Python's max() over (ucb, move, child) tuples (mcts.py:120-123) is a sequential scan that replaces
            // the running best only when the candidate compares greater; a NaN ucb (overflowing logits:
            // exp -> inf, inf/inf) never does.  So: NaN children never win, unless child 0 itself is NaN, in
            // which case nothing ever replaces it.
            double best_u = 0, u0 = 0;
            int best_key = -1, best_i = -1;
            Node best_ch = nd, ch0 = nd;
            for (int i = lane; i < nd.n_children; i += 32) {
                const Node ch = sp.nodes[nd.first_child + i];
                double u;
                if (f64) {
                    double t = __dmul_rn(c, sp.root_p64[(size_t)tree * MAX_MOVES + i]);
                    t = __dmul_rn(t, sq);
                    t = __ddiv_rn(t, (double)(ch.N + 1));
                    u = __dadd_rn((double)ch.Q, t);
                } else {
                    float t = __fmul_rn((float)c, ch.P);
                    t = __fmul_rn(t, (float)sq);
                    t = __fdiv_rn(t, (float)(ch.N + 1));
                    u = (double)__fadd_rn(ch.Q, t);  // f32 value, exactly representable in f64
                }
                if (i == 0) { u0 = u; ch0 = ch; }
                if (u != u) continue;
                const int key = uci_sort_key(ch.move);
                if (best_i < 0 || u > best_u || (u == best_u && key > best_key)) { best_u = u; best_key = key; best_i = i; best_ch = ch; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ou = __shfl_xor_sync(0xffffffffu, best_u, o);
                const int ok = __shfl_xor_sync(0xffffffffu, best_key, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                if (oi >= 0 && (best_i < 0 || ou > best_u || (ou == best_u && ok > best_key))) { best_u = ou; best_key = ok; best_i = oi; }
            }
            u0 = __shfl_sync(0xffffffffu, u0, 0);
            best_i = __shfl_sync(0xffffffffu, best_i, 0);   // one decision for the whole warp
            if (u0 != u0 || best_i < 0) { best_i = 0; best_ch = ch0; }
            node = nd.first_child + best_i;
            // the winning child was read by lane best_i % 32 (child 0 by lane 0): broadcast its entry as the next header
            const int owner = best_i & 31;
            Node nx;
            nx.N = __shfl_sync(0xffffffffu, best_ch.N, owner);
            nx.W = 0.f; nx.Q = 0.f; nx.P = 0.f;
            nx.first_child = __shfl_sync(0xffffffffu, best_ch.first_child, owner);
            nx.parent = 0;
            nx.pos_id = __shfl_sync(0xffffffffu, best_ch.pos_id, owner);
            const int packed = __shfl_sync(0xffffffffu, (int)best_ch.move | ((int)best_ch.n_children << 16) | ((int)(uint8_t)best_ch.term << 24), owner);
            nx.move = (uint16_t)(packed & 0xFFFF);
            nx.n_children = (uint8_t)((packed >> 16) & 0xFF);
            nx.term = (int8_t)(packed >> 24);
            nd = nx;
        }
        __syncwarp();
        if (nd.term == 3) break;  // ran into a leaf that already waits for its evaluation: nothing new to select this step
        // ---- leaf: position, outcome (mcts.py:67, game.py:49-59)
        // lane 0 makes the move; the legal moves come from the whole warp (rules_warp.cuh); the threefold claim-ahead
        // scan (one make-move + history walk per legal move) runs one move per lane
        int term = 0, n_moves = 0, leaf_pos = -1, err = 0, scan = 0, fresh = 0;
        uint16_t* sm = s_m[warp];
        if (lane == 0) {
            term = nd.term;
            leaf_pos = nd.pos_id;
            if (leaf_pos < 0) {
                const Pos pp = sp.pos[parent_pos];  // private copy: the rules code reads its fields many times
                Pos np_ = push_move(pp, nd.move);
                np_.pad_[0] = (uint32_t)parent_pos;
                leaf_pos = chunk_alloc(ts.pos_next, ts.pos_end, 1, POS_CHUNK, sp.pos_count, sp.max_pos);
                if (leaf_pos < 0) {
                    err = 1;
                } else {
                    sp.pos[leaf_pos] = np_;  // occur = 1 for now: the count follows from the whole warp
                    sp.nodes[node].pos_id = leaf_pos;
                    fresh = 1;
                    __threadfence_block();
                }
            }
        }
        __syncwarp();
        fresh = __shfl_sync(0xffffffffu, fresh, 0);
        if (fresh) {
            // first visit of this node: repetition history, legal moves, outcome - all by the whole warp
            leaf_pos = __shfl_sync(0xffffffffu, leaf_pos, 0);
            Pos cur = sp.pos[leaf_pos];  // private copy per lane
            const int W = depth > 0 ? gather_history_warp(sp.pos, sp.nodes, path, depth, ts.root_pos, ts.rec_start, cur.irrev_in != 0,
                                                          s_hh[warp], s_hp[warp], lane)
                                    : -1;
            __syncwarp();
            int occ;
            if (W >= 0) {
                occ = count_occurrences_warp(cur, sp.pos, s_hh[warp], s_hp[warp], W, lane);
            } else {  // path not recorded or a window beyond MCTS_HIST plies: the serial walk
                occ = lane == 0 ? count_occurrences(cur, sp.pos + parent_pos, PoolPrev{sp.pos}) : 0;
                occ = __shfl_sync(0xffffffffu, occ, 0);
            }
            cur.occur = (uint8_t)occ;
            if (lane == 0 && occ != 1) sp.pos[leaf_pos].occur = (uint8_t)occ;
            n_moves = generate_legal_warp(cur, sm, lane);
            __syncwarp();
            if (lane == 0) {
                int out;
                if (W >= 0) {
                    out = outcome_before_history(cur, sm, n_moves, ts.max_len);
                    scan = out == OUT_NONE && W >= 3;  // the claim-ahead scan needs three stored plies behind reversible moves
                } else {
                    bool need_scan = false;
                    out = outcome_before_scan(cur, sm, n_moves, sp.pos + parent_pos, PoolPrev{sp.pos}, ts.max_len, &need_scan);
                    scan = need_scan;
                }
                if (out != OUT_NONE) term = terminal_value_to_play(out) < 0 ? 2 : 1;
            }
            scan = __shfl_sync(0xffffffffu, scan, 0);
            if (scan) {
                bool found = false;
                for (int i = lane; i < n_moves && !found; i += 32)
                    found = W >= 0 ? move_claims_threefold_hist(cur, sm[i], sp.pos, s_hh[warp], s_hp[warp], W)
                                   : move_claims_threefold(cur, sm[i], sp.pos + parent_pos, PoolPrev{sp.pos});
                if (__any_sync(0xffffffffu, found)) term = 1;  // draw by threefold claim
            }
        }
        err = __shfl_sync(0xffffffffu, err, 0);
        term = __shfl_sync(0xffffffffu, term, 0);
        leaf_pos = __shfl_sync(0xffffffffu, leaf_pos, 0);
        if (err) {
            if (lane == 0) { ts.error = 2; atomicAdd(&sp.counters[2], 1); }
            return;
        }
        if (term) {
            // evaluate() returned (value, True): update(node, -value) with value = -1 or 0; terminal leaves are
            // re-evaluated on every visit and never expanded (mcts.py:67-74)
            if (depth > 0) backup_path(sp.nodes, path, depth, term == 2 ? 1.0f : -0.0f, lane);
            __syncwarp();
            if (lane == 0) {
                sp.nodes[node].term = (int8_t)term;
                if (depth <= 0) backup(sp.nodes, node, term == 2 ? 1.0f : -0.0f);
                ts.sims_done += 1;
                atomicAdd(&sp.counters[1], 1);
                __threadfence_block();
            }
            __syncwarp();
            ++sims_done;
            // A simulation that ends in a terminal node needs no evaluation, so the tree goes straight on to the next one - but a tree whose
            // favourite line ends in a claimable draw or a mate would run hundreds of descents inside ONE launch while the other 1023
            // warps have long finished (measured: 50 us typical, 350-500 us spikes, all of it this tail).  After a few terminal
            // simulations the tree yields: it has no leaf in the next batch and carries on in the next step.  The order of its
            // simulations, hence every visit count, is unchanged; every step still advances every unfinished tree by >= 1 simulation.
            if (++n_term >= MCTS_TERMINAL_SIMS_PER_STEP) break;
            continue;
        }
        // ---- a leaf for the network: evaluation slot n_new of this tree
        const int slot = tree * L + n_new;
        {
            // keep the path for the backup that follows the evaluation (phase A of the next step)
            int32_t* gp = sp.path + (size_t)slot * MCTS_MAX_PATH;
            for (int d = lane; d < depth; d += 32) gp[d] = path[d];
            uint16_t* pm = sp.pend_moves + (size_t)slot * MAX_MOVES;
            for (int i = lane; i < n_moves; i += 32) pm[i] = sm[i];
        }
        if (vl && depth > 0) virtual_loss_path(sp.nodes, path, depth, lane);
        __syncwarp();
        int board = 0;
        if (lane == 0) {
            SlotState sl;
            sl.leaf_node = node; sl.leaf_pos = leaf_pos; sl.n_moves = n_moves; sl.path_len = depth > 0 ? depth : 0;
            sp.slots[slot] = sl;
            if (vl) sp.nodes[node].term = 3;
            fill_frames(sp, slot, leaf_pos);
            board = atomicAdd(sp.active_count, 1);
            sp.active_list[board] = slot;  // this leaf is in the next step's evaluation batch
            if (sp.slot_board) sp.slot_board[slot] = board;
            __threadfence_block();
        }
        __syncwarp();
        if (sp.planes_act) {  // fused step: the input planes of the leaf, at its place in the next batch
            board = __shfl_sync(0xffffffffu, board, 0);
            encode_masks_warp(sp.pos, sp.frame_idx + (size_t)slot * 8, s_mask[warp], s_scal[warp], lane);
            encode_cells_warp(s_mask[warp], s_scal[warp], board, sp.planes_act, lane);
            __syncwarp();
        }
        ++n_new;
        if (vl && depth <= 0) break;  // no recorded path, no virtual loss: do not select a second leaf on top of it
    }
    if (lane == 0) {
        ts.n_pending = n_new;
        ts.status = (n_new > 0 || sims_done < sims_target) ? TREE_WAIT_LEAF : TREE_DONE;  // n_new == 0 with simulations left: yielded
    }
}

// Root set-up (mcts.py:50-52): node 0 of each tree with N = 1, its legal moves pending for expand.
__global__ void __launch_bounds__(128) mcts_begin_kernel(SearchParams sp) {
    const int tree = blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= sp.n_trees) return;
    TreeState& ts = sp.trees[tree];
    if (ts.status != TREE_WAIT_ROOT) return;
    const int node = chunk_alloc(ts.node_next, ts.node_end, 1, NODE_CHUNK, sp.node_count, sp.max_nodes);
    if (node < 0) { ts.error = 1; atomicAdd(&sp.counters[2], 1); return; }
    Node r;
    r.N = 1; r.W = 0.f; r.Q = 0.f; r.P = 0.f;
    r.first_child = -1; r.parent = -1; r.pos_id = ts.root_pos; r.move = 0; r.n_children = 0; r.term = 0;
    sp.nodes[node] = r;
    ts.root_node = node;
    ts.n_nodes = 1;
    ts.n_evals = 0;
    ts.sims_done = 0;
    ts.n_pending = 1;
    const int slot0 = tree * sp.leaves;
    MoveList ml;
    generate_legal(sp.pos[ts.root_pos], ml);
    SlotState sl;
    sl.leaf_node = node; sl.leaf_pos = ts.root_pos; sl.n_moves = ml.n; sl.path_len = 0;
    sp.slots[slot0] = sl;
    uint16_t* pm = sp.pend_moves + (size_t)slot0 * MAX_MOVES;
    for (int i = 0; i < ml.n; ++i) pm[i] = ml.m[i];
    for (int i = 0; i < sp.leaves; ++i) fill_frames(sp, slot0 + i, ts.root_pos);  // every slot of the batch encodes a valid position
    const int board = atomicAdd(sp.active_count, 1);
    sp.active_list[board] = slot0;                                                // the root is the first evaluation of the tree
    if (sp.slot_board) sp.slot_board[slot0] = board;
}

int launch_mcts_begin(const SearchParams& sp, cudaStream_t stream) {
    mcts_begin_kernel<<<(sp.n_trees + 127) / 128, 128, 0, stream>>>(sp);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_mcts_step(const SearchParams& sp, cudaStream_t stream) {
    mcts_step_kernel<<<(sp.n_trees + MCTS_WARPS - 1) / MCTS_WARPS, MCTS_WARPS * 32, 0, stream>>>(sp);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------ stand-alone K5
// encode+forward+mask/softmax of BASELINE config 2: one warp per position generates the ordered legal
// moves of pos[cur_idx[i]] and writes moves, action indices and priors.
__global__ void __launch_bounds__(128) policy_softmax_kernel(const Pos* __restrict__ pos, const int32_t* __restrict__ cur_idx, int n,
                                                             const __nv_bfloat16* __restrict__ logits /*[n][4672]*/,
                                                             const float* __restrict__ exp_table, uint16_t* __restrict__ out_moves,
                                                             int16_t* __restrict__ out_actions, float* __restrict__ out_priors,
                                                             int32_t* __restrict__ out_n) {
    __shared__ float s_p[4][MAX_MOVES];
    __shared__ uint16_t s_m[4][MAX_MOVES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + warp;
    if (i >= n) return;
    const Pos p = pos[cur_idx[i]];
    const int nm = generate_legal_warp(p, s_m[warp], lane);
    __syncwarp();
    for (int k = lane; k < nm; k += 32) {
        const int a = move_to_action(s_m[warp][k], p.turn);
        out_moves[(size_t)i * MAX_MOVES + k] = s_m[warp][k];
        out_actions[(size_t)i * MAX_MOVES + k] = (int16_t)a;
        s_p[warp][k] = exp_table[__bfloat16_as_ushort(logits[(size_t)i * 4672 + a])];
    }
    __syncwarp();
    const float sum = numpy_sum_f32_warp(s_p[warp], nm, lane);
    for (int k = lane; k < nm; k += 32) out_priors[(size_t)i * MAX_MOVES + k] = __fdiv_rn(s_p[warp][k], sum);
    if (lane == 0) out_n[i] = nm;
}
// BASELINE config 2 without the dense 4672-logit matrix: one warp per position generates the ordered legal moves of pos[cur_idx[i]] and
// evaluates the policy Dense only for them - a lane per move, the 128-long fmaf chain over the transposed weights WT[action] in the
// same order as the dense kernel and the search's expansion (same bits) - then bf16 -> exp table, NumPy pairwise sum, divide
// (mcts.py:91-101).  ~35 of 4672 logits per position: 130 x less Dense work, no [n][4672] tensor in HBM.
__global__ void __launch_bounds__(128) leaf_priors_kernel(const Pos* __restrict__ pos, const int32_t* __restrict__ cur_idx, int n,
                                                          const float* __restrict__ pfeat, const float* __restrict__ wpt, const float* __restrict__ exp_table,
                                                          uint16_t* __restrict__ out_moves, int16_t* __restrict__ out_actions,
                                                          float* __restrict__ out_priors, int32_t* __restrict__ out_n) {
    __shared__ float s_p[4][MAX_MOVES];
    __shared__ uint16_t s_m[4][MAX_MOVES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + warp;
    if (i >= n) return;
    const Pos p = pos[cur_idx[i]];
    const int nm = generate_legal_warp(p, s_m[warp], lane);
    __syncwarp();
    const float4* pf = reinterpret_cast<const float4*>(pfeat + (size_t)i * 128);
    for (int k = lane; k < nm; k += 32) {
        const int a = move_to_action(s_m[warp][k], p.turn);
        const float4* wr = reinterpret_cast<const float4*>(wpt + (size_t)a * 128);
        float acc = 0.f;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) {
            const float4 w = wr[q], f = pf[q];
            acc = fmaf(f.x, w.x, acc); acc = fmaf(f.y, w.y, acc); acc = fmaf(f.z, w.z, acc); acc = fmaf(f.w, w.w, acc);
        }
        out_moves[(size_t)i * MAX_MOVES + k] = s_m[warp][k];
        out_actions[(size_t)i * MAX_MOVES + k] = (int16_t)a;
        s_p[warp][k] = exp_table[__bfloat16_as_ushort(__float2bfloat16_rn(acc))];
    }
    __syncwarp();
    const float sum = numpy_sum_f32_warp(s_p[warp], nm, lane);
    for (int k = lane; k < nm; k += 32) out_priors[(size_t)i * MAX_MOVES + k] = __fdiv_rn(s_p[warp][k], sum);
    if (lane == 0) out_n[i] = nm;
}
int launch_leaf_priors(const Pos* pos, const int32_t* cur_idx, int n, const float* pfeat, const float* wpt, const float* exp_table,
                       uint16_t* out_moves, int16_t* out_actions, float* out_priors, int32_t* out_n, cudaStream_t stream) {
    if (n <= 0) return 0;
    leaf_priors_kernel<<<(n + 3) / 4, 128, 0, stream>>>(pos, cur_idx, n, pfeat, wpt, exp_table, out_moves, out_actions, out_priors, out_n);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

// dense [n][MAX_MOVES] rows -> one packed buffer (read back with one device-to-host copy): a single block scans the counts, then every
// position's row is copied to its offset.  Sections, contiguous: value f32[n] | n_moves i32[n] | offsets i32[n + 1] | priors f32[T] |
// moves u16[T] | actions i16[T], T = offsets[n].
__global__ void __launch_bounds__(1024) pack_leaf_eval_kernel(int n, const float* __restrict__ value, const int32_t* __restrict__ cnt,
                                                              const uint16_t* __restrict__ moves, const int16_t* __restrict__ actions,
                                                              const float* __restrict__ priors, uint8_t* __restrict__ out) {
    __shared__ int32_t s_scan[1024];
    __shared__ int32_t s_base;
    float* o_value = reinterpret_cast<float*>(out);
    int32_t* o_cnt = reinterpret_cast<int32_t*>(o_value + n);
    int32_t* o_off = o_cnt + n;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int c = i < n ? cnt[i] : 0;
        s_scan[threadIdx.x] = c;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        if (i < n) { o_value[i] = value[i]; o_cnt[i] = c; o_off[i] = s_base + s_scan[threadIdx.x] - c; }
        __syncthreads();
        if (threadIdx.x == 1023) s_base += s_scan[1023];
        __syncthreads();
    }
    const int T = s_base;
    if (threadIdx.x == 0) o_off[n] = T;
    __syncthreads();
    float* o_pri = reinterpret_cast<float*>(o_off + n + 1);
    uint16_t* o_mv = reinterpret_cast<uint16_t*>(o_pri + T);
    int16_t* o_act = reinterpret_cast<int16_t*>(o_mv + T);
    // a warp per position: coalesced row copies
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < n; i += 32) {
        const int c = o_cnt[i], off = o_off[i];
        for (int k = lane; k < c; k += 32) {
            o_pri[off + k] = priors[(size_t)i * MAX_MOVES + k];
            o_mv[off + k] = moves[(size_t)i * MAX_MOVES + k];
            o_act[off + k] = actions[(size_t)i * MAX_MOVES + k];
        }
    }
}
int launch_pack_leaf_eval(int n, const float* value, const int32_t* cnt, const uint16_t* moves, const int16_t* actions, const float* priors,
                          uint8_t* out, cudaStream_t stream) {
    pack_leaf_eval_kernel<<<1, 1024, 0, stream>>>(n, value, cnt, moves, actions, priors, out);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_policy_softmax(const Pos* pos, const int32_t* cur_idx, int n, const __nv_bfloat16* logits, const float* exp_table,
                          uint16_t* out_moves, int16_t* out_actions, float* out_priors, int32_t* out_n, cudaStream_t stream) {
    if (n <= 0) return 0;
    policy_softmax_kernel<<<(n + 3) / 4, 128, 0, stream>>>(pos, cur_idx, n, logits, exp_table, out_moves, out_actions, out_priors, out_n);
    CB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cb
