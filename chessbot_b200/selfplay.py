"""Callers of the hot path: the self-play game loop and replay format of train.py
(train.py:13-52,77-78,104-110) and the puzzle bot of eval_trt.py (eval_trt.py:31-39), over the batched
device search.  Many games advance together: every move of every live game is one tree of one
`mcts.search_batch` call."""
import numpy as np

from . import actionspace as asp
from . import gameimage, mcts
from .config import Config
from .game import Game


def calc_search_statistics(config: Config, root, to_play):
    """train.py:13-19: visit distribution over the 4672 actions, float64."""
    sum_visits = sum([child.N for child in root.children.values()])
    child_visits = np.zeros(config.num_actions)
    for uci_move, child in root.children.items():
        child_visits[asp.uci_to_action(uci_move, to_play)] = child.N / sum_visits
    return child_visits


_ACTION_LUT = {}


def _action_lut(white):
    """action index of every packed move code (from | to << 6 | promo << 12) for one side to move: actionspace.uci_to_action
    (actionspace.py:116-180) as a table, -1 where the reference returns None"""
    lut = _ACTION_LUT.get(bool(white))
    if lut is None:
        from ._lib import lib
        lut = np.array([lib.cb_move_to_action(code, int(bool(white))) for code in range(1 << 15)], dtype=np.int32)
        _ACTION_LUT[bool(white)] = lut
    return lut


def search_statistics_from_root(config: Config, root, to_play):
    """calc_search_statistics for a device root without building its children: the read-back move codes and visit counts go
    through the action table in one gather (same float64 values: N / sum N)."""
    stats = getattr(root, "_stats", None)
    if stats is None:
        return calc_search_statistics(config, root, to_play)
    codes, counts = stats[0], stats[1]
    child_visits = np.zeros(config.num_actions)
    sum_visits = sum(counts.tolist())
    child_visits[_action_lut(to_play)[codes]] = np.array([n / sum_visits for n in counts.tolist()])
    return child_visits


def playout_cap(config: Config, u: float):
    """train.py:29-31 for one uniform draw u: (do_full_search, num_simulations, CPUCT).  The reference compares
    the BOOL with the probability (`do_full_search < num_mcts_sims_p`), so a full search gets num_mcts_sims[1]
    (300) simulations and a quick one num_mcts_sims[0] (50)."""
    do_full_search = True if u < config.num_mcts_sims_p else False
    num_simulations = config.num_mcts_sims[0] if do_full_search < config.num_mcts_sims_p else config.num_mcts_sims[1]
    return do_full_search, num_simulations, (None if do_full_search else 1.0)


def play_games(network, n_games=1, config: Config = None, rngs=None, games=None, lanes=None):
    """train.py:21-52 for n_games games at once.  rngs: one np.random.Generator per game (default: unseeded,
    like the reference); game i consumes its own stream in the reference's order — playout-cap draw, root noise
    (full searches), move choice — so a game does not depend on which other games share its batch.
    lanes: on a DeviceNetwork the games are split into that many groups (default 2 from 64 games up) whose searches alternate on
    the device through mcts.SearchStream: while one group is being searched the host records samples, chooses and plays the
    moves of the other and exports its next positions.  Results do not depend on `lanes`.
    Returns, per game, the reference's play_game tuple:
    (images, (terminal_values, search_statistics), outcome_str, history_len)."""
    import torch
    config = Config() if config is None else config
    games = [Game(config) for _ in range(n_games)] if games is None else list(games)
    n_games = len(games)
    rngs = [np.random.default_rng() for _ in range(n_games)] if rngs is None else list(rngs)
    stats = [[] for _ in range(n_games)]
    images = [[] for _ in range(n_games)]
    on_turn = [[] for _ in range(n_games)]
    device_net = isinstance(network, mcts.DeviceNetwork)
    if lanes is None:
        lanes = 2 if device_net and n_games >= 64 else 1
    lanes = max(1, lanes if device_net else 1)
    stream = mcts.SearchStream(network, lanes) if lanes > 1 else None
    side = torch.cuda.Stream() if stream is not None else None      # sample images are encoded beside the queued searches, not behind them
    groups = [[i for i in range(g, n_games, lanes) if not games[i].terminal()] for g in range(lanes)]

    def submit(live):
        """playout-cap draw and root noise of every live game of a group, then the group's search (queued, or run to completion)"""
        full, sims, cpuct, noise = [], [], [], []
        for i in live:
            f, s, c = playout_cap(config, rngs[i].random())
            full.append(f); sims.append(s); cpuct.append(c)
            noise.append(rngs[i].gamma(mcts.ROOT_DIRICHLET_ALPHA, 1, len(games[i].board.legal_moves)) if f else None)
        batch = [games[i] for i in live]
        if stream is None:
            return full, mcts.search_batch(batch, network, sims, full, cpuct, noise=noise)
        return full, stream.submit(batch, sims, full, cpuct, noise=noise)

    def finish(live, full, roots):
        """samples of the full searches (train.py:43-47), move choice, moves played; returns the games still running"""
        rec = [k for k, f in enumerate(full) if f]
        if rec:
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())       # nothing to wait for in practice: orders allocator reuse
                with torch.cuda.stream(side):
                    imgs = gameimage.boards_to_images([games[live[k]].board for k in rec], config.T).astype(np.int16)
            else:
                imgs = gameimage.boards_to_images([games[live[k]].board for k in rec], config.T).astype(np.int16)
            for img, k in zip(imgs, rec):
                g = games[live[k]]
                stats[live[k]].append(search_statistics_from_root(config, roots[k], g.to_play()))
                images[live[k]].append(img)
                on_turn[live[k]].append(g.to_play())
        for k, i in enumerate(live):
            temp = mcts.TEMP if games[i].history_len < config.num_mcts_sampling_moves else 0
            root = roots[k]                    # device roots: move choice from the read-back counts (same NumPy / RNG calls as select_move)
            games[i].make_move(mcts._select_move_from_counts(root._stats[0], root._stats[1], rngs[i], temp) if hasattr(root, "_stats")
                               else mcts.select_move(root, rngs[i], temp))
        return [i for i in live if not games[i].terminal()]

    pending = [submit(g) if g else None for g in groups]
    while any(p is not None for p in pending):
        for gi in range(lanes):
            if pending[gi] is None:
                continue
            full, res = pending[gi]
            roots = res if stream is None else stream.collect(res)
            groups[gi] = finish(groups[gi], full, roots)
            pending[gi] = submit(groups[gi]) if groups[gi] else None
    out = []
    for i, g in enumerate(games):
        terminal_values = [g.terminal_value(player) for player in on_turn[i]]
        out.append((images[i], (terminal_values, stats[i]), g.outcome_str, g.history_len))
    return out


def play_game(network, config: Config = None, rng=None):
    """train.py:21-52, one game."""
    return play_games(network, 1, config, None if rng is None else [rng])[0]


def buffer_entries(game_result):
    """train.py:77-78: the (image, (terminal_value, search_statistic)) samples a finished game appends to the buffer."""
    images, (terminal_values, search_statistics), _, _ = game_result
    return [(img, (z, pi)) for img, z, pi in zip(images, terminal_values, search_statistics)]


def sample_buffer(buffer, batch_size: int, rng=None):
    """train.py:104-110: uniform sampling without replacement."""
    rng = np.random.default_rng() if rng is None else rng
    batch = rng.choice(len(buffer), size=batch_size, replace=False)
    images = np.array([buffer[b][0] for b in batch])
    terminal_values = np.array([buffer[b][1][0] for b in batch])
    search_statistics = np.array([buffer[b][1][1] for b in batch])
    return images, (terminal_values, search_statistics)


class Bot:
    """eval_trt.py:31-39 (`bot.get_move(board)` is all puzzles.Puzzle.solve needs, puzzles.py:83-86), plus
    get_moves for many boards in one device batch.  Deterministic by default: no root noise, arg-max visits."""

    def __init__(self, model, config: Config = None, num_simulations=None, add_noise=False, rng=None):
        self.model = model
        self.config = Config() if config is None else config
        self.num_simulations = self.config.num_mcts_sims[1] if num_simulations is None else num_simulations
        self.add_noise = add_noise
        self.rng = np.random.default_rng() if rng is None else rng

    def get_moves(self, boards):
        games = [Game(self.config, board=b) for b in boards]
        res = mcts.run_mcts_batch(games, self.model, self.num_simulations, 0, self.add_noise, None, rng=self.rng)
        return [move for move, _ in res]

    def get_move(self, board):
        return self.get_moves([board])[0]


def solve_puzzles(puzzles, bot, max_rounds=64):
    """puzzles.PuzzleSet.solve / eval_trt.solve_puzzles (puzzles.py:83-86,195-215; eval_trt.py:41-52) for many puzzles at once:
    every round, the current positions of all puzzles still IN_PROGRESS are searched as ONE device batch (`bot.get_moves`) and
    each answer is played with the puzzle's own `move_uci` (which also plays the opponent's scripted reply).  `puzzles` are
    the reference's `Puzzle` objects, unmodified (anything with `.board`, `.status`, `.move_uci(uci)`); the result per puzzle is
    what `puzzle.solve(bot)` leaves behind.  Returns the number of search rounds."""
    rounds = 0
    live = [p for p in puzzles if p.status == "IN_PROGRESS"]
    while live and rounds < max_rounds:
        for p, move in zip(live, bot.get_moves([p.board for p in live])):
            p.move_uci(move)
        live = [p for p in live if p.status == "IN_PROGRESS"]
        rounds += 1
    return rounds


def run_actors(jobs, device=None):
    """Several actors on ONE GPU, the way the reference deploys them (num_actors processes per GPU, train.py:173,198-201),
    as threads of this process: jobs[i] is a callable taking the actor's own `engine.Engine` and returning its result.  Every actor thread gets its own device context and CUDA stream, so
    while one actor waits for its searches (the C ABI releases the GIL) the others do their host work and keep the GPU fed.
    Returns the results in job order; the first exception is re-raised.
    Measured on one B200, 20x256: no gain.  Two actors of 1024 games each search at the same evals/s as one (a single actor's
    queue of steps already keeps the GPU busy, and two towers evict each other's 49 MB of weights from L2); in playout-cap
    self-play 2 actors x 1024 games reach 315 k sims/s against 386 k for ONE actor with 2048 games (scripts/selfplay_soak.py).
    More games per actor is the better lever; actors remain useful when each has too few games to fill the tensor cores."""
    import threading

    import torch

    from .engine import Engine
    dev = torch.cuda.current_device() if device is None else int(device)
    results, errors = [None] * len(jobs), []

    def work(i):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                results[i] = jobs[i](Engine.actor(i, dev))
                torch.cuda.current_stream().synchronize()
        except BaseException as e:       # noqa: BLE001 - handed to the caller
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results
