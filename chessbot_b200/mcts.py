"""Drop-in for the reference's mcts.py (mcts.py:22-161).

`run_mcts` keeps the reference's signature and return value but the search itself — PUCT select,
device chess rules, expand (legal-move gather, exp, normalise) and backup — runs in the CUDA node
pool of csrc/mcts.cu.  `run_mcts_batch` is the same call for many games at once (one tree per game,
one leaf per tree per step: every tree keeps the reference's sequential simulation order, so visit
counts are those of the reference given the same network outputs and noise).

The node-level helpers (`expand`, `evaluate`, `update`, `select_leaf`, `ucb`, `select_move`,
`add_exploration_noise`) keep working on Python `Node` objects for callers that use them directly
(tests/mcts_v2.py); they follow the reference's NumPy-2 float semantics with a float32 network.
"""
import math

import numpy as np

from . import actionspace as asp
from .game import Game
from .model import DeviceNetwork, SyntheticNetwork, predict_fn

TEMP = 1.0
PB_C_BASE = 19652
PB_C_INIT = 1.25
ROOT_DIRICHLET_ALPHA = 0.3
ROOT_EXPLORATION_FRACTION = 0.25


class Node:
    __slots__ = ("parent", "children", "N", "W", "P", "Q", "__dict__")

    def __init__(self):
        self.parent = None
        self.children = {}
        self.N = 0
        self.W = 0
        self.P = 0
        self.Q = 0

    def is_leaf(self):
        return len(self.children) == 0


# ------------------------------------------------------------------------------ device search
_UCI = {}


def _uci(code):
    """uci string of a packed move (from | to << 6 | promo << 12), memoised: 1792 x 5 distinct values at most."""
    u = _UCI.get(code)
    if u is None:
        from .board import code_to_uci
        u = _UCI[code] = code_to_uci(code)
    return u


class _DeviceRoot(Node):
    """Root of a device search.  N is set at once; the children dict (legal-move order, one Node per child with the reference's
    scalar types: np.float32 W / Q / P, np.float64 P at a noisy root) is built from the read-back arrays the first time
    `children` is touched - a caller that only wants the move (puzzles, quick playout-cap searches) never pays for ~30 Python
    objects per tree."""

    def __init__(self, root_n, codes, ns, ws, qs, ps):
        self.parent = None
        self.N, self.W, self.P, self.Q = root_n, 0, 0, 0
        self._stats = (codes, ns, ws, qs, ps)

    @property
    def children(self):
        d = self.__dict__.get("_children")
        if d is None:
            codes, ns, ws, qs, ps = self._stats
            d, new = {}, Node.__new__
            for u, n_, w_, q_, p_ in zip(map(_uci, codes.tolist()), ns.tolist(), list(ws), list(qs), list(ps)):
                c = new(Node)
                c.parent, c.children, c.N, c.W, c.P, c.Q = self, {}, n_, w_, p_, q_
                d[u] = c
            self.__dict__["_children"] = d
        return d

    @children.setter
    def children(self, value):
        self.__dict__["_children"] = value


def _roots_from_device(out, n):
    """Root Nodes of n trees over views of the read-back root statistics (children built lazily, see _DeviceRoot).  `out` is either the
    packed read-back (flat arrays + "off", engine.search_read_roots_compact) or the dense [n, 224] one (engine.search_read_roots)."""
    k = out["n_children"][:n].tolist()
    root_n = out["root_N"][:n].tolist()
    moves, N, W, Q, P, noisy = out["moves"], out["N"], out["W"], out["Q"], out["P"], out["noisy"]
    p32 = None if all(noisy[:n]) else P.astype(np.float32)          # priors are float32 unless root noise made them float64
    if "off" in out:
        lo = out["off"][:n].tolist()
        return [_DeviceRoot(root_n[i], moves[a:a + c], N[a:a + c], W[a:a + c], Q[a:a + c], P[a:a + c] if noisy[i] else p32[a:a + c])
                for i, (a, c) in enumerate(zip(lo, k))]
    return [_DeviceRoot(root_n[i], moves[i, :k[i]], N[i, :k[i]], W[i, :k[i]], Q[i, :k[i]], P[i, :k[i]] if noisy[i] else p32[i, :k[i]])
            for i in range(n)]


def _draw_move(counts, rng, temp):
    """Index of the move select_move (mcts.py:142-153) picks for these visit counts: the same NumPy / Generator calls."""
    visit_counts = np.asarray(counts).astype(np.int64)
    if temp == 0:
        return int(rng.choice(np.flatnonzero(visit_counts == np.max(visit_counts))))
    pi = visit_counts ** (1 / temp)
    pi = pi / np.sum(pi)
    return int(np.where(rng.multinomial(1, pi) == 1)[0][0])


def _select_move_from_counts(codes, counts, rng, temp):
    """select_move on a device root's arrays, without building the children"""
    return _uci(int(codes[_draw_move(counts, rng, temp)]))


def _select_moves(roots, temps, rng, flat=None):
    """select_move for every root of a batch.  A greedy choice (temp 0) with a single most-visited child draws nothing from the
    generator (Generator.choice over one candidate consumes no bits), so those - the bulk of a batch - are one vectorised arg-max
    over a padded matrix; the games whose choice does consume randomness (temp 1, or ties at temp 0) make the reference's own calls,
    in game order.  flat = (codes, counts, off, k) of the packed read-back the roots are views of (else gathered from the roots)."""
    n = len(roots)
    if n == 0:
        return []
    if flat is None:
        k = np.fromiter((len(r._stats[1]) for r in roots), dtype=np.int64, count=n)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(k, out=off[1:])
        codes = np.concatenate([r._stats[0] for r in roots])
        counts = np.concatenate([r._stats[1] for r in roots])
    else:
        codes, counts, off, k = flat
    col = np.arange(max(int(k.max()), 1))
    inside = col[None, :] < np.asarray(k)[:, None]
    cm = np.where(inside, counts[np.minimum(np.asarray(off[:n], dtype=np.int64)[:, None] + col[None, :], len(counts) - 1)], -1)
    is_top = cm == cm.max(1)[:, None]
    pick = is_top.argmax(1)
    temps = np.asarray(temps)
    for i in np.flatnonzero((temps != 0) | (is_top.sum(1) > 1)).tolist():
        pick[i] = _draw_move(counts[off[i]:off[i] + k[i]], rng, temps[i])
    return [_uci(c) for c in codes[np.asarray(off[:n], dtype=np.int64) + pick].tolist()]


def _external_search(engine, network, records, sims, cpuct, noise, max_len):
    from .engine import EVAL_EXTERNAL
    engine.search_begin(records, sims, cpuct=cpuct, noise=noise, max_game_length=max_len, eval_mode=EVAL_EXTERNAL)
    n = len(records)
    values = np.zeros(n, dtype=np.float32)
    expvals = np.zeros((n, 4672), dtype=np.float32)
    while True:
        planes, waiting = engine.search_pending_planes()
        if not waiting.any():
            break
        for i in np.flatnonzero(waiting):
            value, logits = predict_fn(network, planes[i].astype(np.float32))
            values[i] = np.float32(np.asarray(value, dtype=np.float32))
            with np.errstate(over="ignore"):
                expvals[i] = np.exp(np.asarray(logits, dtype=np.float32), dtype=np.float32)  # mcts.py:94
        engine.search_step_external(values, expvals)
    st = engine.search_status()
    if st["errors"]:
        raise RuntimeError("search: node or position pool exhausted")
    out = engine.search_read_roots()
    out.update(st)
    return out


def _prepare_batch(games, num_simulations, add_noise, CPUCT, rng, noise):
    """Host side of a batched search before the device is involved: game records (one C call), per-game settings, root noise in CSR
    form.  num_simulations / add_noise / CPUCT may be scalars or per-game lists."""
    from .board import export_records
    from .gameimage import as_device_board
    n = len(games)
    sims = [num_simulations] * n if np.isscalar(num_simulations) else list(num_simulations)
    noisy = [bool(add_noise)] * n if isinstance(add_noise, (bool, np.bool_)) else [bool(a) for a in add_noise]
    cpuct = [CPUCT] * n if (CPUCT is None or np.isscalar(CPUCT)) else list(CPUCT)
    rec_flat, rec_lens, n_legal = export_records([as_device_board(g.board) for g in games], flat=True)
    noff = np.zeros(n + 1, dtype=np.int64)
    if noise is None:
        # one draw for all noisy roots: Generator.gamma fills sequentially, so the ranges are the per-game draws of mcts.py:158
        rng = np.random.default_rng() if rng is None else rng
        np.cumsum(np.where(np.asarray(noisy, dtype=bool), n_legal, 0), out=noff[1:])
        flat = rng.gamma(ROOT_DIRICHLET_ALPHA, 1, int(noff[-1])) if noff[-1] else np.zeros(0)
        noisy_flags = noisy
    else:
        # caller-drawn samples: cut / zero-pad every noisy root's row to its legal-move count (the device indexes them by child)
        noisy_flags = [bool(f) and z is not None for z, f in zip(noise, noisy)]
        np.cumsum(np.where(np.asarray(noisy_flags, dtype=bool), n_legal, 0), out=noff[1:])
        flat = np.zeros(int(noff[-1]))
        for i in np.flatnonzero(noisy_flags).tolist():
            z = np.asarray(noise[i], dtype=np.float64)[:n_legal[i]]
            flat[noff[i]:noff[i] + len(z)] = z
    return {"n": n, "records": (rec_flat, rec_lens), "n_legal": n_legal, "sims": sims, "cpuct": cpuct,
            "noise": (flat, noff) if noff[-1] else None, "noisy": noisy_flags, "max_len": games[0].config.max_game_length}


def _ensure_pool(engine, n, max_sims):
    shape = engine.search_shape
    if shape is None or shape[0] < n or shape[3] < max_sims:
        trees = max(n, shape[0] if shape else 0)
        ms = max(max_sims, shape[3] if shape else 0)
        engine.search_create(trees, max_nodes=min(trees * (ms + 2) * 48, 400_000_000), max_positions=trees * (ms + 200), max_sims=ms)


def _roots_of(out, prep):
    out["noisy"] = prep["noisy"]
    roots = _roots_from_device(out, prep["n"])
    if "off" in out and roots:           # lets run_mcts_batch choose all moves from the packed arrays in one pass
        roots[0]._batch_flat = (out["moves"], out["N"], out["off"], out["n_children"][:prep["n"]])
    return roots


def search_batch(games, network, num_simulations=100, add_noise=True, CPUCT=None, rng=None, noise=None, leaves_per_tree=1):
    """The search of run_mcts (mcts.py:50-74) for a list of games in one device batch, without the final move
    choice.  num_simulations / add_noise / CPUCT may be scalars or per-game lists.  `noise`: optional per-game
    pre-drawn Gamma(0.3, 1) samples (else drawn from `rng` like mcts.py:158).  leaves_per_tree > 1 selects that many
    leaves of every tree per step under virtual loss: more throughput for few games, visit counts no longer those of
    the reference's sequential order (device networks only).  Returns the root Nodes."""
    from .engine import EVAL_NET, EVAL_SYNTH, Engine
    prep = _prepare_batch(games, num_simulations, add_noise, CPUCT, rng, noise)
    n, sims = prep["n"], prep["sims"]
    engine = network.engine if isinstance(network, (DeviceNetwork, SyntheticNetwork)) else Engine.get()
    _ensure_pool(engine, n, max(max(sims), 1))
    kw = dict(cpuct=prep["cpuct"], noise=prep["noise"], max_game_length=prep["max_len"])
    if isinstance(network, DeviceNetwork):
        network.ensure_loaded()
        if network.max_batch < n * leaves_per_tree:
            network.max_batch = n * leaves_per_tree
            network.load()
        out = engine.search(prep["records"], sims, compact_counts=prep["n_legal"], eval_mode=EVAL_NET, leaves=leaves_per_tree, **kw)
    elif isinstance(network, SyntheticNetwork):
        out = engine.search(prep["records"], sims, compact_counts=prep["n_legal"], eval_mode=EVAL_SYNTH, seed=network.seed, leaves=leaves_per_tree, **kw)
    else:
        if leaves_per_tree != 1:
            raise ValueError("a host network evaluates one leaf per tree: leaves_per_tree must be 1")
        out = _external_search(engine, network, prep["records"], sims, prep["cpuct"], prep["noise"], prep["max_len"])
    return _roots_of(out, prep)


class SearchStream:
    """Several batched searches in flight on one GPU (the self-play / puzzle drivers' way to keep the device busy while the host
    chooses moves, plays them and exports the next records).  Every lane is a device context of its own (node pool, activations, a
    copy of the network: model.DeviceNetwork.sibling); all lanes queue on the caller's current CUDA stream, so the searches run one
    after the other in submission order, and collect() waits for ITS search's read-back event only - the next search is already
    running.  Results are those of search_batch (a tree's visit counts do not depend on what else is in flight).

        stream = SearchStream(network)            # two lanes
        t_a = stream.submit(games_a, 100); t_b = stream.submit(games_b, 100)
        roots_a = stream.collect(t_a)             # device is busy with games_b meanwhile
    """

    def __init__(self, network, lanes=2):
        if not isinstance(network, DeviceNetwork):
            raise TypeError("SearchStream runs on a DeviceNetwork (a host callable needs the device between steps)")
        self.networks = [network] + [network.sibling(i) for i in range(1, lanes)]
        self._busy = [False] * lanes
        self._next = 0

    def submit(self, games, num_simulations=100, add_noise=True, CPUCT=None, rng=None, noise=None):
        """Exports the games, queues the whole search (root expansion + max(num_simulations) steps) and its packed read-back; returns a
        ticket at once.  At most one search per lane: collect() the oldest ticket before submitting lane-count + 1 searches."""
        from .engine import EVAL_NET
        free = [l for l in range(len(self.networks)) if not self._busy[l]]
        if not free:
            raise RuntimeError("SearchStream: every lane holds a search that was not collected")
        lane = self._next if self._next in free else free[0]
        prep = _prepare_batch(games, num_simulations, add_noise, CPUCT, rng, noise)
        net = self.networks[lane]
        engine, n, sims = net.engine, prep["n"], prep["sims"]
        _ensure_pool(engine, n, max(max(sims), 1))
        net.ensure_loaded()
        if net.max_batch < n:
            net.max_batch = n
            net.load()
        engine.search_begin(prep["records"], sims, cpuct=prep["cpuct"], noise=prep["noise"], max_game_length=prep["max_len"], eval_mode=EVAL_NET,
                            leaves=1)
        engine.search_run(int(max(sims)) + 1)
        prep["ticket"] = engine.search_read_roots_compact_async(prep["n_legal"])
        self._busy[lane] = True
        self._next = (lane + 1) % len(self.networks)
        return lane, prep

    def collect(self, ticket):
        """Root Nodes of a submitted search (blocks until its read-back has landed)."""
        lane, prep = ticket
        out = self.networks[lane].engine.search_read_roots_compact_wait(prep["ticket"])
        self._busy[lane] = False
        if out["errors"]:
            raise RuntimeError("search: node or position pool exhausted (raise max_nodes / max_positions)")
        if out["unfinished"]:
            raise RuntimeError("search did not finish")
        self.last_counters = {k: out[k] for k in ("evals", "sims")}
        return _roots_of(out, prep)

    def collect_moves(self, ticket, games, num_sampling_moves=30, rng=None):
        """[(move_uci, root)] like run_mcts_batch for a submitted search."""
        roots = self.collect(ticket)
        return _moves_for(games, roots, num_sampling_moves, np.random.default_rng() if rng is None else rng)


def _moves_for(games, roots, num_sampling_moves, rng):
    temps = [TEMP if g.history_len < num_sampling_moves else 0 for g in games]
    return list(zip(_select_moves(roots, temps, rng, getattr(roots[0], "_batch_flat", None) if roots else None), roots))


def run_mcts_batch(games, network, num_simulations=100, num_sampling_moves=30, add_noise=True, CPUCT=None,
                   rng=None, noise=None, leaves_per_tree=1):
    """run_mcts for a list of games in one device batch (one tree per game).  Returns [(move_uci, root Node), ...]."""
    rng = np.random.default_rng() if rng is None else rng
    roots = search_batch(games, network, num_simulations, add_noise, CPUCT, rng=rng, noise=noise, leaves_per_tree=leaves_per_tree)
    return _moves_for(games, roots, num_sampling_moves, rng)


def run_mcts(game: Game, network, num_simulations: int = 100, num_sampling_moves=30, add_noise=True, CPUCT=None, leaves_per_tree=1):
    """mcts.py:38-77: returns (move, root) with root.children[uci].N/W/Q/P filled in.  leaves_per_tree (extension, default 1 =
    the reference's behaviour): leaves evaluated per network call under virtual loss."""
    return run_mcts_batch([game], network, num_simulations, num_sampling_moves, add_noise, CPUCT, leaves_per_tree=leaves_per_tree)[0]


# ------------------------------------------------------------------------------ node-level API
# The reference's helpers on Python Node objects (tests/mcts_v2.py drives them directly).  Signatures, argument meaning and float
# semantics are the reference's (NumPy 2 weak scalars: float32 priors / W / Q / ucb with a float32 network, float64 at a noisy root);
# the bodies are written over this package's primitives: move list and action indices come from the C-ABI rules engine in one call
# each, the exponentials and their sum are one vectorised NumPy expression (np.sum over a float32 array is the same pairwise sum the
# reference gets from its list of float32 scalars).
def expand(node: Node, game: Game, network=None):
    """mcts.py:80-103: evaluates the position, attaches one child per legal move (legal-move order) with prior
    exp(logit) / sum over the legal moves, returns the value as a 0-d float32 array."""
    value, logits = predict_fn(network, game.make_image(-1).astype(np.float32))
    ucis = game.legal_moves()
    actions = np.fromiter((asp.uci_to_action(u, game.to_play()) for u in ucis), dtype=np.int64, count=len(ucis))
    weights = np.exp(np.asarray(logits, dtype=np.float32)[actions], dtype=np.float32)
    priors = weights / np.sum(weights, dtype=np.float32)
    for uci, prior in zip(ucis, priors):            # iterating a float32 array yields np.float32 scalars, like the reference's dict
        child = Node()
        child.parent, child.P = node, prior
        node.children[uci] = child
    return np.array(value, dtype=np.float32)


def evaluate(game: Game):
    """mcts.py:106-110: (value for the side to move, True) at a terminal position, else (0, False)."""
    over = bool(game.terminal())
    return (game.terminal_value(game.to_play()) if over else 0), over


def update(node: Node, value: float):
    """mcts.py:112-117: one more visit worth `value` at node, -value at its parent, ... up to the root."""
    sign = 1
    while node is not None:
        node.N += 1
        node.W += value if sign > 0 else -value
        node.Q = node.W / node.N
        node, sign = node.parent, -sign


def _exploration_rate(parent_visits):
    return math.log((parent_visits + PB_C_BASE + 1) / PB_C_BASE) + PB_C_INIT


def ucb(parent: Node, child: Node, CPUCT=None):
    """mcts.py:132-140: Q + c * P * sqrt(N_parent) / (N_child + 1), evaluated left to right (its rounding order is part of the
    search's results, see csrc/mcts.cu)."""
    c = _exploration_rate(parent.N) if CPUCT is None else CPUCT
    return child.Q + c * child.P * parent.N ** 0.5 / (child.N + 1)


def select_leaf(node: Node, CPUCT=None):
    """mcts.py:119-130: (move, child) of the highest ucb; equal scores go to the lexicographically larger move string."""
    best = None
    for move, child in node.children.items():
        key = (ucb(node, child, CPUCT), move)
        if best is None or key > best[0]:           # the comparisons max() makes over the reference's tuples, NaN behaviour included
            best = (key, child)
    return best[0][1], best[1]


def select_move(node: Node, rng: np.random.Generator = None, temp=0):
    """mcts.py:142-153: the most visited move (ties drawn uniformly), or one drawn in proportion to N ** (1 / temp)."""
    rng = np.random.default_rng() if rng is None else rng
    moves = list(node.children)
    return moves[_draw_move([child.N for child in node.children.values()], rng, temp)]


def add_exploration_noise(node: Node, rng: np.random.Generator):
    """mcts.py:156-161: P <- 0.75 P + 0.25 Gamma(0.3, 1) per child, in children order (un-normalised, as the reference has it)."""
    keep = 1 - ROOT_EXPLORATION_FRACTION
    for child, z in zip(node.children.values(), rng.gamma(ROOT_DIRICHLET_ALPHA, 1, len(node.children))):
        child.P = child.P * keep + z * ROOT_EXPLORATION_FRACTION
