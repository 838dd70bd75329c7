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
    """Root Nodes of n trees over views of the read-back root statistics (children built lazily, see _DeviceRoot)."""
    k = out["n_children"][:n].tolist()
    root_n = out["root_N"][:n].tolist()
    moves, N, W, Q, P, noisy = out["moves"], out["N"], out["W"], out["Q"], out["P"], out["noisy"]
    p32 = None if all(noisy[:n]) else P[:n].astype(np.float32)      # priors are float32 unless root noise made them float64
    return [_DeviceRoot(root_n[i], moves[i, :k[i]], N[i, :k[i]], W[i, :k[i]], Q[i, :k[i]], P[i, :k[i]] if noisy[i] else p32[i, :k[i]])
            for i in range(n)]


def _select_move_from_counts(codes, counts, rng, temp):
    """select_move (mcts.py:142-153) on a device root's arrays: the same NumPy / RNG calls, without building the children"""
    visit_counts = counts.astype(np.int64)
    if temp == 0:
        return _uci(int(codes[rng.choice(np.flatnonzero(visit_counts == np.max(visit_counts)))]))
    pi = visit_counts ** (1 / temp)
    pi = pi / np.sum(pi)
    return _uci(int(codes[np.where(rng.multinomial(1, pi) == 1)[0][0]]))


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


def search_batch(games, network, num_simulations=100, add_noise=True, CPUCT=None, rng=None, noise=None, leaves_per_tree=1):
    """The search of run_mcts (mcts.py:50-74) for a list of games in one device batch, without the final move
    choice.  num_simulations / add_noise / CPUCT may be scalars or per-game lists.  `noise`: optional per-game
    pre-drawn Gamma(0.3, 1) samples (else drawn from `rng` like mcts.py:158).  leaves_per_tree > 1 selects that many
    leaves of every tree per step under virtual loss: more throughput for few games, visit counts no longer those of
    the reference's sequential order (device networks only).  Returns the root Nodes."""
    from .engine import EVAL_NET, EVAL_SYNTH, Engine
    from .gameimage import as_device_board
    n = len(games)
    sims = [num_simulations] * n if np.isscalar(num_simulations) else list(num_simulations)
    noisy = [bool(add_noise)] * n if isinstance(add_noise, (bool, np.bool_)) else [bool(a) for a in add_noise]
    cpuct = [CPUCT] * n if (CPUCT is None or np.isscalar(CPUCT)) else list(CPUCT)
    from .board import export_records
    boards = [as_device_board(g.board) for g in games]
    records, n_legal = export_records(boards)
    if noise is None:
        # one draw for all noisy roots: Generator.gamma fills sequentially, so the slices are the per-game draws of mcts.py:158
        rng = np.random.default_rng() if rng is None else rng
        counts = [int(k) if z else 0 for k, z in zip(n_legal, noisy)]
        flat = rng.gamma(ROOT_DIRICHLET_ALPHA, 1, sum(counts)) if any(counts) else np.zeros(0)
        ends = np.cumsum(counts)
        noise = [flat[e - k:e] if z else None for e, k, z in zip(ends.tolist(), counts, noisy)]
    else:
        noise = [z if f else None for z, f in zip(noise, noisy)]
    max_len = games[0].config.max_game_length
    engine = network.engine if isinstance(network, (DeviceNetwork, SyntheticNetwork)) else Engine.get()
    max_sims = max(max(sims), 1)
    shape = engine.search_shape
    if shape is None or shape[0] < n or shape[3] < max_sims:
        trees = max(n, shape[0] if shape else 0)
        ms = max(max_sims, shape[3] if shape else 0)
        engine.search_create(trees, max_nodes=min(trees * (ms + 2) * 48, 400_000_000), max_positions=trees * (ms + 200), max_sims=ms)
    has_noise = any(z is not None for z in noise)
    if isinstance(network, DeviceNetwork):
        network.ensure_loaded()
        if network.max_batch < n * leaves_per_tree:
            network.max_batch = n * leaves_per_tree
            network.load()
        out = engine.search(records, sims, cpuct=cpuct, noise=noise if has_noise else None, max_game_length=max_len, eval_mode=EVAL_NET,
                            leaves=leaves_per_tree)
    elif isinstance(network, SyntheticNetwork):
        out = engine.search(records, sims, cpuct=cpuct, noise=noise if has_noise else None, max_game_length=max_len,
                            eval_mode=EVAL_SYNTH, seed=network.seed, leaves=leaves_per_tree)
    else:
        if leaves_per_tree != 1:
            raise ValueError("a host network evaluates one leaf per tree: leaves_per_tree must be 1")
        out = _external_search(engine, network, records, sims, cpuct, noise if has_noise else None, max_len)
    out["noisy"] = [z is not None for z in noise]
    return _roots_from_device(out, n)


def run_mcts_batch(games, network, num_simulations=100, num_sampling_moves=30, add_noise=True, CPUCT=None,
                   rng=None, noise=None, leaves_per_tree=1):
    """run_mcts for a list of games in one device batch (one tree per game).  Returns [(move_uci, root Node), ...]."""
    rng = np.random.default_rng() if rng is None else rng
    roots = search_batch(games, network, num_simulations, add_noise, CPUCT, rng=rng, noise=noise, leaves_per_tree=leaves_per_tree)
    results = []
    for g, root in zip(games, roots):
        temp = TEMP if g.history_len < num_sampling_moves else 0
        codes, counts = root._stats[0], root._stats[1]
        results.append((_select_move_from_counts(codes, counts, rng, temp), root))
    return results


def run_mcts(game: Game, network, num_simulations: int = 100, num_sampling_moves=30, add_noise=True, CPUCT=None, leaves_per_tree=1):
    """mcts.py:38-77: returns (move, root) with root.children[uci].N/W/Q/P filled in.  leaves_per_tree (extension, default 1 =
    the reference's behaviour): leaves evaluated per network call under virtual loss."""
    return run_mcts_batch([game], network, num_simulations, num_sampling_moves, add_noise, CPUCT, leaves_per_tree=leaves_per_tree)[0]


# ------------------------------------------------------------------------------ node-level API
def expand(node: Node, game: Game, network=None):
    """mcts.py:80-103 for one Python node: value from the network, priors for the legal moves."""
    value, policy_logits = predict_fn(network, game.make_image(-1).astype(np.float32))
    value = np.array(value, dtype=np.float32)
    policy_logits = np.array(policy_logits, dtype=np.float32)
    policy = {}
    for move in game.legal_moves():
        action = asp.uci_to_action(move, game.to_play())
        policy[move] = np.exp(policy_logits[action], dtype=np.float32)
    policy_sum = np.sum(list(policy.values()), dtype=np.float32)
    for move, p in policy.items():
        child = Node()
        child.parent = node
        child.P = p / policy_sum
        node.children[move] = child
    return value


def evaluate(game: Game):
    if game.terminal():
        return game.terminal_value(game.to_play()), True
    return 0, False


def update(node: Node, value: float):
    while node:
        node.N += 1
        node.W += value
        node.Q = node.W / node.N
        value = -value
        node = node.parent


def ucb(parent: Node, child: Node, CPUCT=None):
    if CPUCT is None:
        CPUCT = math.log((parent.N + PB_C_BASE + 1) / PB_C_BASE) + PB_C_INIT
    return child.Q + CPUCT * child.P * parent.N ** 0.5 / (child.N + 1)


def select_leaf(node: Node, CPUCT=None):
    _, move, child = max((ucb(node, child, CPUCT), move, child) for move, child in node.children.items())
    return move, child


def select_move(node: Node, rng: np.random.Generator = None, temp=0):
    rng = np.random.default_rng() if rng is None else rng
    moves = [move for move in node.children.keys()]
    visit_counts = [child.N for child in node.children.values()]
    if temp == 0:
        return moves[rng.choice(np.flatnonzero(visit_counts == np.max(visit_counts)))]
    pi = np.array(visit_counts) ** (1 / temp)
    pi = pi / np.sum(pi)
    return moves[np.where(rng.multinomial(1, pi) == 1)[0][0]]


def add_exploration_noise(node: Node, rng: np.random.Generator):
    moves = node.children.keys()
    noise = rng.gamma(ROOT_DIRICHLET_ALPHA, 1, len(moves))
    frac = ROOT_EXPLORATION_FRACTION
    for i, move in enumerate(moves):
        node.children[move].P = node.children[move].P * (1 - frac) + noise[i] * frac
