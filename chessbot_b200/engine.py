"""Python host of the B200 hot path: owns device memory through PyTorch and calls the C ABI.

PyTorch is plumbing here (allocation, streams); every computation is a hand-written sm_100a kernel in
libchessbot_b200.so.  Nothing in this module falls back to the CPU: without a B200 `Engine()` raises.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import CB_MAX_MOVES, CB_NUM_ACTIONS, CB_POS_BYTES, check, last_error, lib

EVAL_NET, EVAL_SYNTH, EVAL_EXTERNAL = 0, 1, 2
_engines = {}

# ---- device memory: the library allocates what it keeps between calls through PyTorch's caching allocator (cb_set_device_allocator)
_live_blocks = {}


@C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)
def _torch_alloc(nbytes, device, _user):
    try:
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=torch.device("cuda", int(device)))
    except Exception:                      # noqa: BLE001 - reported by the library as an allocation failure
        return None
    _live_blocks[t.data_ptr()] = t
    return t.data_ptr()


@C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
def _torch_free(ptr, _user):
    _live_blocks.pop(ptr, None)


def library_device_bytes():
    """bytes of device memory the library currently holds (all of it PyTorch tensors)"""
    return sum(t.numel() for t in _live_blocks.values())


_allocator_installed = False


def _install_allocator():
    global _allocator_installed
    if not _allocator_installed:
        check(lib.cb_set_device_allocator(_torch_alloc, _torch_free, None), "set_device_allocator")
        _allocator_installed = True
        import atexit
        atexit.register(lambda: lib.cb_set_device_allocator(None, None, None))   # no callbacks into a finalising interpreter


def _stream(device=None):
    """torch's current stream of `device` (default: the current device) as the `void* stream` of the C ABI"""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def numpy_exp_table():
    """f32 exp for every bf16 bit pattern, from this host's np.exp (mcts.py:94 uses np.exp(float32))."""
    bits = (np.arange(65536, dtype=np.uint32) << 16).view(np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        return np.exp(bits, dtype=np.float32)


def act_to_nhwc(act, n, channels):
    """tile-native activation buffer (bf16) -> [n, 8, 8, channels] (test / debug helper)."""
    pairs = (n + 1) // 2
    v = act.view(torch.bfloat16)[: pairs * (channels // 8) * 1600].view(pairs, channels // 8, 10, 2, 10, 8)
    v = v[:, :, 1:9, :, 1:9, :]                       # P, j, row, b, col, e
    return v.permute(0, 3, 2, 4, 1, 5).reshape(pairs * 2, 8, 8, channels)[:n]


def nhwc_to_act(x):
    """[n, 8, 8, C] (C multiple of 8) -> tile-native bf16 buffer with zero border."""
    n, _, _, c = x.shape
    pairs = (n + 1) // 2
    buf = torch.zeros(pairs, c // 8, 10, 2, 10, 8, dtype=torch.bfloat16, device=x.device)
    xp = torch.zeros(pairs * 2, 8, 8, c, dtype=torch.bfloat16, device=x.device)
    xp[:n] = x.to(torch.bfloat16)
    buf[:, :, 1:9, :, 1:9, :] = xp.view(pairs, 2, 8, 8, c // 8, 8).permute(0, 4, 2, 1, 3, 5)
    return buf.reshape(-1)


class Engine:
    """One device context per GPU (cb_ctx)."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("chessbot_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device_index)
        _install_allocator()
        self._h = lib.cb_ctx_create(self.device_index)
        if not self._h:
            raise RuntimeError(f"cb_ctx_create: {last_error()}")
        self.sm_count = lib.cb_ctx_sm_count(self._h)
        tab = numpy_exp_table()
        check(lib.cb_set_exp_table(self._h, tab.ctypes.data), "set_exp_table")
        self.net_shape = None
        self.search_shape = None

    @classmethod
    def get(cls, device=None):
        idx = torch.cuda.current_device() if device is None else int(device)
        if idx not in _engines:
            _engines[idx] = cls(idx)
        return _engines[idx]

    @classmethod
    def actor(cls, index, device=None):
        """Device context number `index` of extra actors sharing this GPU (kept for the life of the process; index 0, 1, ...).
        Every context owns its network copy, node pool and scratch, so actors on different CUDA streams do not interfere."""
        idx = torch.cuda.current_device() if device is None else int(device)
        key = (idx, "actor", int(index))
        if key not in _engines:
            _engines[key] = cls(idx)
        return _engines[key]

    # ------------------------------------------------------------------ gameimage
    def upload_records(self, records):
        """records: list of uint8 [k_i, 96] game records (Board.export_record) -> (pos tensor, frame_idx [n,8])."""
        if isinstance(records, tuple):
            flat, lens = np.ascontiguousarray(records[0]), np.asarray(records[1], dtype=np.int64)
        else:
            lens = np.array([len(r) for r in records], dtype=np.int64)
            flat = np.concatenate(records, axis=0)
        # frame lists, newest first: back to the start of the record or of the game (ply 0), at most 8 (gameimage.py:170-181)
        ply = flat[:, 74].astype(np.int64) | (flat[:, 75].astype(np.int64) << 8)
        last = np.cumsum(lens) - 1
        n_frames = np.minimum(np.minimum(lens, 8), ply[last] + 1)
        t = np.arange(8)
        frame_idx = np.where(t[None, :] < n_frames[:, None], last[:, None] - t[None, :], -1).astype(np.int32)
        pos = torch.from_numpy(flat).to(self.device)
        return pos, torch.from_numpy(frame_idx).to(self.device)

    def encode(self, pos, frame_idx, want_act=True, want_i16=False):
        n = frame_idx.shape[0]
        act = torch.zeros(lib.cb_planes_act_bytes(n) // 2, dtype=torch.bfloat16, device=self.device) if want_act else None
        i16 = torch.empty(n, 8, 8, 119, dtype=torch.int16, device=self.device) if want_i16 else None
        check(lib.cb_encode(self._h, _ptr(pos), _ptr(frame_idx), n, _ptr(act), _ptr(i16), _stream(self.device_index)), "encode")
        return act, i16

    # ------------------------------------------------------------------ model
    def load_network(self, weights, filters, blocks, max_batch, slope=0.3, eps=1e-3, stream_hi_lo=True):
        """stream_hi_lo: residual stream kept as bf16 hi + lo between blocks (cb_net_set_stream_precision)"""
        check(lib.cb_net_create(self._h, filters, blocks, max_batch, slope, eps), "net_create")
        # True / False, or an int k >= 2: hi + lo only from residual block k - 1 on
        mode = int(stream_hi_lo) if isinstance(stream_hi_lo, (bool, np.bool_)) else int(stream_hi_lo)
        check(lib.cb_net_set_stream_precision(self._h, mode), "net_set_stream_precision")
        for name, arr in weights.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            check(lib.cb_net_set_tensor(self._h, name.encode(), a.ctypes.data, a.size), "net_set_tensor")
        check(lib.cb_net_finalize(self._h), "net_finalize")
        self.net_shape = (filters, blocks, max_batch)

    def forward(self, planes_act, n, want_bf16=True, want_f32=False):
        value = torch.empty(n, dtype=torch.float32, device=self.device)
        lb = torch.empty(n, CB_NUM_ACTIONS, dtype=torch.bfloat16, device=self.device) if want_bf16 else None
        lf = torch.empty(n, CB_NUM_ACTIONS, dtype=torch.float32, device=self.device) if want_f32 else None
        check(lib.cb_net_forward(self._h, _ptr(planes_act), n, _ptr(value), _ptr(lb), _ptr(lf), _stream(self.device_index)), "net_forward")
        return value, lb, lf

    def synth_forward(self, planes_i16, seed=0):
        n = planes_i16.shape[0]
        value = torch.empty(n, dtype=torch.float32, device=self.device)
        lb = torch.empty(n, CB_NUM_ACTIONS, dtype=torch.bfloat16, device=self.device)
        check(lib.cb_synth_forward(self._h, _ptr(planes_i16), n, seed, _ptr(value), _ptr(lb), _stream(self.device_index)), "synth_forward")
        return value, lb

    def policy_softmax(self, pos, cur_idx, logits_bf16):
        n = cur_idx.shape[0]
        moves = torch.zeros(n, CB_MAX_MOVES, dtype=torch.int16, device=self.device)
        actions = torch.zeros(n, CB_MAX_MOVES, dtype=torch.int16, device=self.device)
        priors = torch.zeros(n, CB_MAX_MOVES, dtype=torch.float32, device=self.device)
        counts = torch.zeros(n, dtype=torch.int32, device=self.device)
        check(lib.cb_policy_softmax(self._h, _ptr(pos), _ptr(cur_idx), n, _ptr(logits_bf16), _ptr(moves), _ptr(actions),
                                    _ptr(priors), _ptr(counts), _stream(self.device_index)), "policy_softmax")
        return moves, actions, priors, counts

    def evaluate_positions(self, records, pos=None, frame_idx=None, dense_logits=False):
        """BASELINE config 2 — batched leaf evaluation outside a tree: encode (K1) -> tower + heads (K2/K3) -> legal-move gather / exp /
        normalise (K5).  records: game records of the positions (host), or pos / frame_idx already on the device.  The policy Dense is
        evaluated for the legal moves only (cb_leaf_eval); dense_logits=True takes the round-1 route through the dense [n, 4672] bf16
        logits (K4 + cb_policy_softmax; same bits).  Returns device tensors (value [n], moves [n,224] uint16 codes as int16, actions
        [n,224], priors [n,224] f32, n_moves [n])."""
        if pos is None:
            pos, frame_idx = self.upload_records(records)
        n = frame_idx.shape[0]
        if dense_logits:
            act, _ = self.encode(pos, frame_idx)
            value, lb, _ = self.forward(act, n, want_bf16=True, want_f32=False)
            moves, actions, priors, counts = self.policy_softmax(pos, frame_idx[:, 0].contiguous(), lb)
            return value, moves, actions, priors, counts
        key = ("leaf_scratch", n)
        if getattr(self, "_leaf_scratch_key", None) != key:
            self._leaf_scratch = torch.zeros(lib.cb_planes_act_bytes(n) // 2, dtype=torch.bfloat16, device=self.device)
            self._leaf_scratch_key = key
        value = torch.empty(n, dtype=torch.float32, device=self.device)
        moves = torch.zeros(n, CB_MAX_MOVES, dtype=torch.int16, device=self.device)
        actions = torch.zeros(n, CB_MAX_MOVES, dtype=torch.int16, device=self.device)
        priors = torch.zeros(n, CB_MAX_MOVES, dtype=torch.float32, device=self.device)
        counts = torch.zeros(n, dtype=torch.int32, device=self.device)
        cur = frame_idx[:, 0].contiguous()
        check(lib.cb_leaf_eval(self._h, _ptr(pos), _ptr(frame_idx), _ptr(cur), n, _ptr(self._leaf_scratch), _ptr(value), _ptr(moves), _ptr(actions),
                               _ptr(priors), _ptr(counts), _stream(self.device_index)), "leaf_eval")
        return value, moves, actions, priors, counts

    def evaluate_records(self, records):
        """The same end to end for HOST game records (list of uint8 [k_i, 96], or (flat, lens) as board.export_records(flat=True)
        returns): one C call - pinned staging, H2D, kernels, packed device-to-host read.  Returns host arrays (views of a buffer the
        engine reuses: copy what must outlive the next call)
        {"value" [n], "n_moves" [n], "off" [n+1], "priors" [T], "moves" [T] uint16, "actions" [T] int16}: position i owns [off[i], off[i+1])."""
        if isinstance(records, tuple):
            flat, lens = np.ascontiguousarray(records[0]), np.ascontiguousarray(records[1], dtype=np.int32)
        else:
            lens = np.array([len(r) for r in records], dtype=np.int32)
            flat = np.ascontiguousarray(np.concatenate(records, axis=0))
        n = len(lens)
        buf = getattr(self, "_leaf_out", None)                  # reused between calls (the views returned are valid until the next call)
        if buf is None or buf.nbytes < lib.cb_leaf_eval_packed_bytes(n, n * CB_MAX_MOVES):
            buf = self._leaf_out = np.zeros(lib.cb_leaf_eval_packed_bytes(n, n * CB_MAX_MOVES), dtype=np.uint8)
        T = check(lib.cb_leaf_eval_host_sync(self._h, flat.ctypes.data, lens.ctypes.data, n, buf.ctypes.data, buf.nbytes, _stream(self.device_index)), "leaf_eval_host")
        o, out = 0, {}
        for key, dt, cnt in (("value", np.float32, n), ("n_moves", np.int32, n), ("off", np.int32, n + 1), ("priors", np.float32, T),
                             ("moves", np.uint16, T), ("actions", np.int16, T)):
            nb = cnt * np.dtype(dt).itemsize
            out[key] = buf[o:o + nb].view(dt)
            o += nb
        return out

    PROF_STAGES = ("encode", "conv", "heads", "mcts_step", "policy_dense", "policy_softmax")

    def profile(self, enable):
        """CUDA events around every kernel launch of the entry points (measurement legs of bench.py)."""
        check(lib.cb_profile(self._h, int(enable)), "profile")

    def profile_read(self):
        """{stage: (total_ms, launches)} since profile(True) / the last read."""
        ms = (C.c_double * len(self.PROF_STAGES))()
        n = (C.c_int * len(self.PROF_STAGES))()
        check(lib.cb_profile_read_sync(self._h, ms, n), "profile_read")
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PROF_STAGES)}

    def debug_conv_layer(self, layer, in_act, skip_act, n_boards, out_channels, use_reference=False):
        out = torch.zeros(((n_boards + 1) // 2) * (out_channels // 8) * 1600, dtype=torch.bfloat16, device=self.device)
        check(lib.cb_debug_conv_layer(self._h, layer, _ptr(in_act), _ptr(skip_act), _ptr(out), n_boards, int(use_reference),
                                      _stream(self.device_index)), "debug_conv_layer")
        return out

    def debug_wgrad(self, in_act, dy_act, n_boards, cin_pad, n, use_reference=False):
        out = torch.zeros(9, cin_pad, n, dtype=torch.float32, device=self.device)
        check(lib.cb_debug_wgrad(self._h, _ptr(in_act), _ptr(dy_act), n_boards, cin_pad, n, _ptr(out), int(use_reference), _stream(self.device_index)),
              "debug_wgrad")
        return out

    # ------------------------------------------------------------------ mcts
    def search_create(self, max_trees, max_nodes=None, max_positions=None, max_sims=800):
        max_nodes = max_nodes or max_trees * (max_sims + 2) * 40
        max_positions = max_positions or max_trees * (max_sims + 2 + 160)
        if self.search_shape == (max_trees, max_nodes, max_positions, max_sims):
            return
        need = lib.cb_search_workspace_bytes(max_trees, max_nodes, max_positions, max_sims, 1)
        free, _total = torch.cuda.mem_get_info(self.device_index)
        if need > free + torch.cuda.memory_reserved(self.device_index) - torch.cuda.memory_allocated(self.device_index):
            raise RuntimeError(f"search pool of {need / 2**30:.1f} GiB (trees {max_trees}, nodes {max_nodes}, positions {max_positions}) does not fit the "
                               f"{free / 2**30:.1f} GiB free on cuda:{self.device_index}: lower max_nodes / max_sims or the number of trees")
        check(lib.cb_search_create(self._h, max_trees, max_nodes, max_positions, max_sims), "search_create")
        self.search_shape = (max_trees, max_nodes, max_positions, max_sims)
        self.search_leaves = 1

    def search_begin(self, records, sims, cpuct=None, noise=None, max_game_length=512, eval_mode=EVAL_NET, seed=0, leaves=1):
        """records: list of uint8 [k_i,96], or (flat uint8 [sum k_i,96], int32 lens) as board.export_records(flat=True) returns them;
        sims: int or list; cpuct: None / float / list (None = AlphaZero schedule); noise: None, a list (per tree) of None / float64
        arrays of Gamma(0.3,1) samples, or CSR (flat float64, int64 offsets [n+1]).
        leaves: leaves of a tree evaluated per step — 1 = the reference's sequential simulations (exact visit
        counts), > 1 = throughput mode with virtual loss (evaluation batch = trees * leaves).
        Does not synchronise with the device (the library stages the inputs in pinned memory)."""
        if isinstance(records, tuple):
            flat, lens = records
            flat, lens = np.ascontiguousarray(flat), np.ascontiguousarray(lens, dtype=np.int32)
        else:
            lens = np.array([len(r) for r in records], dtype=np.int32)
            flat = np.ascontiguousarray(np.concatenate(records, axis=0))
        n = len(lens)
        if leaves != getattr(self, "search_leaves", 1):
            check(lib.cb_search_set_leaves(self._h, int(leaves)), "search_set_leaves")
            self.search_leaves = int(leaves)
        sims_a = np.full(n, sims, dtype=np.int32) if np.isscalar(sims) else np.asarray(sims, dtype=np.int32)
        if cpuct is None or np.isscalar(cpuct):
            cp = np.full(n, -1.0 if cpuct is None else float(cpuct), dtype=np.float64)
        else:
            cp = np.array([-1.0 if c is None else float(c) for c in cpuct], dtype=np.float64)
        nz_flat = nz_off = None
        if isinstance(noise, tuple):
            nz_flat, nz_off = np.ascontiguousarray(noise[0], dtype=np.float64), np.ascontiguousarray(noise[1], dtype=np.int64)
        elif noise is not None:
            parts = [np.zeros(0) if z is None else np.asarray(z, dtype=np.float64)[:CB_MAX_MOVES] for z in noise]
            nz_off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum([len(z) for z in parts], out=nz_off[1:])
            nz_flat = np.ascontiguousarray(np.concatenate(parts)) if nz_off[-1] else None
        self._n_trees = n
        check(lib.cb_search_begin_csr(self._h, n, flat.ctypes.data, lens.ctypes.data, sims_a.ctypes.data, cp.ctypes.data,
                                      nz_flat.ctypes.data if nz_flat is not None and len(nz_flat) else None,
                                      nz_off.ctypes.data if nz_off is not None else None,
                                      max_game_length, eval_mode, seed, _stream(self.device_index)), "search_begin")

    def search_run(self, steps):
        check(lib.cb_search_run(self._h, steps, _stream(self.device_index)), "search_run")

    def search_status(self):
        c = np.zeros(4, dtype=np.int64)
        check(lib.cb_search_status_sync(self._h, c.ctypes.data, _stream(self.device_index)), "search_status")
        return {"unfinished": int(c[0]), "evals": int(c[1]), "sims": int(c[2]), "errors": int(c[3])}

    def search_pending_planes(self):
        n = self._n_trees
        planes = np.empty((n, 8, 8, 119), dtype=np.int16)
        waiting = np.zeros(n, dtype=np.uint8)
        check(lib.cb_search_pending_planes_sync(self._h, planes.ctypes.data, waiting.ctypes.data, _stream(self.device_index)), "pending_planes")
        return planes, waiting.astype(bool)

    def search_step_external(self, values, expvals):
        v = np.ascontiguousarray(values, dtype=np.float32)
        e = np.ascontiguousarray(expvals, dtype=np.float32)
        check(lib.cb_search_step_external(self._h, v.ctypes.data, e.ctypes.data, _stream(self.device_index)), "step_external")

    def search_read_roots(self):
        n = self._n_trees
        moves = np.zeros((n, CB_MAX_MOVES), dtype=np.uint16)
        N = np.zeros((n, CB_MAX_MOVES), dtype=np.int32)
        W = np.zeros((n, CB_MAX_MOVES), dtype=np.float32)
        Q = np.zeros((n, CB_MAX_MOVES), dtype=np.float32)
        P = np.zeros((n, CB_MAX_MOVES), dtype=np.float64)
        nc = np.zeros(n, dtype=np.int32)
        root_n = np.zeros(n, dtype=np.int32)
        nodes = np.zeros(n, dtype=np.int32)
        check(lib.cb_search_read_roots_sync(self._h, moves.ctypes.data, N.ctypes.data, W.ctypes.data, Q.ctypes.data,
                                            P.ctypes.data, nc.ctypes.data, root_n.ctypes.data, nodes.ctypes.data, _stream(self.device_index)),
              "read_roots")
        return {"moves": moves, "N": N, "W": W, "Q": Q, "P": P, "n_children": nc, "root_N": root_n, "nodes": nodes}

    def search_read_roots_compact_async(self, n_legal):
        """Queues the packed read-back of the root statistics behind the steps already on the stream (no synchronisation); returns
        the ticket search_read_roots_compact_wait() takes."""
        n = self._n_trees
        off = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(n_legal, out=off[1:])
        check(lib.cb_search_read_roots_compact_async(self._h, off.ctypes.data, _stream(self.device_index)), "read_roots_compact_async")
        return n, off

    def search_read_roots_compact_wait(self, ticket):
        """Blocks until that read-back has landed (only that: later work on the stream keeps running).  Flat arrays + "off" [n+1]:
        the children of tree t are [off[t], off[t] + n_children[t]); plus the counters of search_status()."""
        n, off = ticket
        T = int(off[-1])
        buf = np.empty(lib.cb_search_roots_compact_bytes(n, T), dtype=np.uint8)
        c = np.zeros(4, dtype=np.int64)
        check(lib.cb_search_read_roots_compact_wait(self._h, buf.ctypes.data, c.ctypes.data), "read_roots_compact_wait")
        o = 0
        out = {"off": off, "unfinished": int(c[0]), "evals": int(c[1]), "sims": int(c[2]), "errors": int(c[3])}
        for key, dt, cnt in (("P", np.float64, T), ("N", np.int32, T), ("W", np.float32, T), ("Q", np.float32, T), ("n_children", np.int32, n),
                             ("root_N", np.int32, n), ("nodes", np.int32, n), ("state", np.int32, n), ("moves", np.uint16, T)):
            nb = cnt * np.dtype(dt).itemsize
            out[key] = buf[o:o + nb].view(dt)
            o += nb
        return out

    def search_read_roots_compact(self, n_legal):
        """Root statistics packed by the roots' legal-move counts: one device-to-host copy."""
        return self.search_read_roots_compact_wait(self.search_read_roots_compact_async(n_legal))

    def search(self, records, sims, compact_counts=None, **kw):
        """Complete batched search on the device network / synthetic network: returns read_roots() + counters
        (compact_counts = the roots' legal-move counts: search_read_roots_compact() instead)."""
        self.search_begin(records, sims, **kw)
        max_sims = int(np.max(sims))
        leaves = self.search_leaves
        self.search_run((max_sims + leaves - 1) // leaves + 1)
        if compact_counts is not None and leaves == 1:
            # sequential simulations finish in exactly max_sims + 1 steps: status and root statistics come back in one copy
            out = self.search_read_roots_compact(compact_counts)
            if out["errors"]:
                raise RuntimeError("search: node or position pool exhausted (raise max_nodes / max_positions)")
            if out["unfinished"]:
                raise RuntimeError("search did not finish")
            return out
        st = self.search_status()
        extra = 0
        while st["unfinished"] and not st["errors"] and leaves > 1 and extra < max_sims:
            self.search_run(4)                       # steps that selected fewer than `leaves` leaves (collisions) leave a remainder
            extra += 4
            st = self.search_status()
        if st["errors"]:
            raise RuntimeError("search: node or position pool exhausted (raise max_nodes / max_positions)")
        if st["unfinished"]:
            raise RuntimeError("search did not finish")
        out = self.search_read_roots() if compact_counts is None else self.search_read_roots_compact(compact_counts)
        out.update(st)
        return out
