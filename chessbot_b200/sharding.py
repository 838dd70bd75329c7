"""Multi-GPU: self-play / puzzle searches are independent per game (train.py:21-52 shares no state between
games; actors only append finished samples, train.py:77-78), so the hot path shards BY GAME with no
data-path collective: game g runs on rank g mod world.  One process per GPU
(torchrun); torch.distributed moves only the small host-side results (moves, visit counts) to rank 0."""
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items, rank, world_size):
    """Indices of the games this rank searches: g -> rank g mod world_size."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_items, world_size))


def sharded_map(fn, items, gather=True, group=None):
    """Applies fn(list_of_items) -> list_of_results to this rank's shard of `items`.  With gather=True
    every rank returns the full result list in the original order (one all_gather_object of the host
    results); with gather=False it returns (indices, results) of its own shard."""
    rank, ws = world()
    idx = shard_indices(len(items), rank, ws)
    local = fn([items[i] for i in idx]) if idx else []
    if len(local) != len(idx):
        raise ValueError("fn must return one result per item")
    if not gather:
        return idx, local
    if ws == 1:
        return list(local)
    parts = [None] * ws
    dist.all_gather_object(parts, (idx, list(local)), group=group)
    out = [None] * len(items)
    for ids, res in parts:
        for i, r in zip(ids, res):
            out[i] = r
    return out


def run_mcts_sharded(games, network, num_simulations=100, num_sampling_moves=30, add_noise=True, CPUCT=None,
                     rng=None, noise=None, gather=True):
    """mcts.run_mcts_batch over all ranks: every rank searches games[rank::world] on its own GPU with its own
    copy of the network.  Per-game lists (num_simulations / add_noise / CPUCT / noise) are sharded with the
    games.  Returns [(move, root)] for all games (gather=True) or (indices, results) of this rank's shard."""
    import numpy as np
    from . import mcts
    rank, ws = world()

    def pick(x, ids, scalar_ok=True):
        if x is None or (scalar_ok and (np.isscalar(x) or isinstance(x, (bool, np.bool_)))):
            return x
        return [x[i] for i in ids]

    def fn(_):
        ids = shard_indices(len(games), rank, ws)
        return mcts.run_mcts_batch([games[i] for i in ids], network, pick(num_simulations, ids), num_sampling_moves,
                                   pick(add_noise, ids), pick(CPUCT, ids), rng=rng, noise=pick(noise, ids, scalar_ok=False))

    return sharded_map(fn, list(range(len(games))), gather=gather)
