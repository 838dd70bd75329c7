"""ORACLE (test infrastructure). Restatement of mcts.py:22-161 as a struct-of-arrays tree.

The reference keeps Python objects and lets NumPy pick dtypes.  Under NumPy 2
(NEP 50; the only stack that can be run beside the GPU) with a network that
returns float32 (TF tensors in the reference, model.py:73-80) that resolves to:

* priors      p = np.exp(logit, dtype=f32); sum = np.sum(list, dtype=f32) (pairwise,
              see `np_sum_f32_model`); P = p / sum in f32           (mcts.py:94-100)
* backup      N += 1; W += v (f32 add); Q = W / N (f32 div), sign flips per ply (mcts.py:112-117)
* ucb         Q + ((f32(C) * P) * f32(N_parent ** 0.5)) / f32(N_child + 1), every op rounded
              to f32, no FMA; C = log((N_parent + 19653) / 19652) + 1.25 in f64 (mcts.py:126-140)
* noisy root  P = f64(f32(P * 0.75f)) + noise_f64 * 0.25 is float64, so root ucb runs in
              f64 with Q upcast                                         (mcts.py:156-161)
* argmax tie  -> lexicographically largest UCI string               (mcts.py:120-123)
* select_move host RNG on visit counts                                (mcts.py:142-153)

Terminal-only nodes hold Python ints in the reference; their W = +-N or 0 and Q = +-1 or 0 are
exact in f32, so f32 storage is equivalent (root.Q differs only if every simulation was terminal).
Pinned against the reference's mcts.py imported verbatim (scripts/make_golden.py -> tests/golden/mcts_*.npz).
"""
import math

import numpy as np

from . import ref_actionspace as asp

TEMP = 1.0
PB_C_BASE = 19652
PB_C_INIT = 1.25
ROOT_DIRICHLET_ALPHA = 0.3
ROOT_EXPLORATION_FRACTION = 0.25
f32 = np.float32
f64 = np.float64


def np_sum_f32_model(a):
    """Explicit model of np.sum(float32[n], dtype=float32): NumPy's pairwise reduction."""
    a = np.asarray(a, dtype=f32)
    n = len(a)

    def block(x):
        m = len(x)
        if m < 8:
            r = f32(0.0) if m == 0 else x[0]
            for v in x[1:]:
                r = f32(r + v)
            return r
        if m <= 128:
            r = [x[j] for j in range(8)]
            i = 8
            while i < m - (m % 8):
                for j in range(8):
                    r[j] = f32(r[j] + x[i + j])
                i += 8
            res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
            while i < m:
                res = f32(res + x[i])
                i += 1
            return res
        h = m // 2
        h -= h % 8
        return f32(block(x[:h]) + block(x[h:]))

    if n == 0:
        return f32(0.0)
    return block(a)


def default_exp(x):
    return np.exp(np.asarray(x, dtype=f32), dtype=f32)


class Tree:
    """Node pool: index 0 is the root. Children of a node are contiguous, in legal-move order."""

    def __init__(self):
        self.parent = [-1]
        self.first_child = [-1]
        self.n_children = [0]
        self.move = [None]
        self.N = [0]
        self.W = [f32(0.0)]
        self.Q = [f32(0.0)]
        self.P = [f32(0.0)]  # f64 entries for noisy-root children

    def add_children(self, node, moves, priors):
        self.first_child[node] = len(self.parent)
        self.n_children[node] = len(moves)
        for m, p in zip(moves, priors):
            self.parent.append(node)
            self.first_child.append(-1)
            self.n_children.append(0)
            self.move.append(m)
            self.N.append(0)
            self.W.append(f32(0.0))
            self.Q.append(f32(0.0))
            self.P.append(p)

    def children(self, node):
        fc = self.first_child[node]
        return range(fc, fc + self.n_children[node]) if fc >= 0 else range(0)

    def is_leaf(self, node):
        return self.n_children[node] == 0

    def root_children(self):
        return {self.move[c]: c for c in self.children(0)}


def legal_actions(game):
    to_play = game.to_play()
    moves = game.legal_moves()
    return moves, [asp.uci_to_action(m, to_play) for m in moves]


def expand(tree, node, game, network, exp_fn=default_exp):
    image = game.make_image(-1).astype(f32)
    out = network(image[None])
    value = f32(np.asarray(out["value_head"], dtype=f32)[0][0])
    logits = np.asarray(out["policy_head"], dtype=f32)[0]
    moves, actions = legal_actions(game)
    p = exp_fn(logits[np.asarray(actions, dtype=np.int64)]) if moves else np.zeros(0, f32)
    s = np.sum(p, dtype=f32)
    tree.add_children(node, moves, [f32(x / s) for x in p])
    return value


def add_exploration_noise(tree, noise):
    frac = ROOT_EXPLORATION_FRACTION
    for i, c in enumerate(tree.children(0)):
        tree.P[c] = f64(f32(tree.P[c] * f32(1 - frac))) + f64(noise[i]) * frac


def cpuct_of(n_parent, cpuct):
    if cpuct is None:
        return math.log((n_parent + PB_C_BASE + 1) / PB_C_BASE) + PB_C_INIT
    return float(cpuct)


def ucb(tree, parent, child, cpuct=None):
    n_parent = tree.N[parent]
    c = cpuct_of(n_parent, cpuct)
    sq = n_parent ** 0.5
    p = tree.P[child]
    if isinstance(p, np.float64):
        u = f64(c) * p
        u = u * f64(sq)
        u = u / f64(tree.N[child] + 1)
        return f64(tree.Q[child]) + u
    u = f32(f32(c) * p)
    u = f32(u * f32(sq))
    u = f32(u / f32(tree.N[child] + 1))
    return f32(tree.Q[child] + u)


def select_child(tree, node, cpuct=None):
    best = None
    for c in tree.children(node):
        key = (ucb(tree, node, c, cpuct), tree.move[c])
        if best is None or key > best[0]:
            best = (key, c)
    return best[1]


def backup(tree, node, value):
    v = f32(value)
    while node >= 0:
        tree.N[node] += 1
        tree.W[node] = f32(tree.W[node] + v)
        tree.Q[node] = f32(tree.W[node] / f32(tree.N[node]))
        v = f32(-v)
        node = tree.parent[node]


def select_move(tree, rng, temp=0):
    kids = list(tree.children(0))
    moves = [tree.move[c] for c in kids]
    counts = np.array([tree.N[c] for c in kids])
    if temp == 0:
        return moves[rng.choice(np.flatnonzero(counts == np.max(counts)))]
    pi = counts ** (1 / temp)
    pi = pi / np.sum(pi)
    return moves[np.where(rng.multinomial(1, pi) == 1)[0][0]]


def run_mcts(game, network, num_simulations=100, num_sampling_moves=30, add_noise=True, CPUCT=None,
             rng=None, noise=None, exp_fn=default_exp, stats=None):
    """mcts.py:38-77.  `noise`: pre-drawn Gamma(0.3, 1) samples for the root (else drawn from rng)."""
    rng = np.random.default_rng() if rng is None else rng
    tree = Tree()
    tree.N[0] = 1
    expand(tree, 0, game, network, exp_fn)
    n_evals = 1
    if add_noise:
        if noise is None:
            noise = rng.gamma(ROOT_DIRICHLET_ALPHA, 1, tree.n_children[0])
        add_exploration_noise(tree, noise)
    for _ in range(num_simulations):
        node = 0
        tmp = game.clone()
        while not tree.is_leaf(node):
            node = select_child(tree, node, CPUCT)
            tmp.make_move(tree.move[node])
        if tmp.terminal():
            value = tmp.terminal_value(tmp.to_play())
        else:
            value = expand(tree, node, tmp, network, exp_fn)
            n_evals += 1
        backup(tree, node, -value)
    if stats is not None:
        stats["evals"] = n_evals
    temp = TEMP if game.history_len < num_sampling_moves else 0
    return select_move(tree, rng, temp), tree
