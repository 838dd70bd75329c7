"""N>1 host logic of the training step on CPU: world_size-2 gloo processes each take their shard of a global batch
(chessbot_b200.train.shard_batch), compute gradients (the oracle stands in for the device kernels, which need a GPU),
sum them with chessbot_b200.train.allreduce_gradients and apply the SGD update: both ranks must end with the same weights,
equal to one process averaging the two shards' gradients.  (BatchNormalization statistics stay per rank, as on the GPUs.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILTERS, BLOCKS, N, LR = 8, 1, 8, 0.05


def _data():
    from oracle import ref_model, ref_train
    w = ref_model.init_weights(FILTERS, BLOCKS, seed=5, random_bn=True)
    rng = np.random.default_rng(5)
    images = rng.integers(0, 2, (N, 8, 8, 119)).astype(np.int16)
    z, pi = ref_train.synthetic_targets(N, 5)
    return w, images, z, pi


def _flat(g, keys):
    return np.concatenate([np.asarray(g[k], np.float32).ravel() for k in keys])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chessbot_b200 import train as cb_train
    from oracle import ref_train
    assert cb_train.data_parallel_world() == (rank, world)
    w, images, z, pi = _data()
    sl = cb_train.shard_batch(N)
    _, g, stats, _, _ = ref_train.forward_backward(w, images[sl], z[sl], pi[sl])
    keys = sorted(w)
    flat = torch.from_numpy(_flat(g, keys))
    scale = cb_train.allreduce_gradients(flat)
    q.put((rank, (sl.start, sl.stop), scale, flat.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_batch_partitions_the_global_batch():
    from chessbot_b200.train import shard_batch
    for n in (1, 7, 64, 4096):
        for w in (1, 2, 3, 8):
            parts = [shard_batch(n, r, w) for r in range(w)]
            assert parts[0].start == 0 and parts[-1].stop == n
            assert all(parts[i].stop == parts[i + 1].start for i in range(w - 1))
            sizes = [p.stop - p.start for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo_gradient_allreduce_matches_one_process():
    from oracle import ref_train
    w, images, z, pi = _data()
    keys = sorted(w)
    halves = [ref_train.forward_backward(w, images[s], z[s], pi[s])[1] for s in (slice(0, N // 2), slice(N // 2, N))]
    expected_sum = _flat(halves[0], keys) + _flat(halves[1], keys)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [(0, N // 2), (N // 2, N)]
    for rank, _, scale, flat in got:
        assert scale == 0.5
        assert np.allclose(flat, expected_sum, rtol=1e-3, atol=1e-4 * float(np.abs(expected_sum).max()))   # the summed gradient (fp32 order differs)
    assert np.array_equal(got[0][3], got[1][3])


def _worker_uneven(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chessbot_b200 import train as cb_train
    from oracle import ref_train
    w, images, z, pi = _data()
    n = N - 1                                            # 7 samples over 2 ranks: shards of 4 and 3
    sl = cb_train.shard_batch(n)
    _, g, _, _, _ = ref_train.forward_backward(w, images[:n][sl], z[:n][sl], pi[:n][sl])
    keys = sorted(w)
    flat = torch.from_numpy(_flat(g, keys))
    scale = cb_train.allreduce_gradients(flat, local_n=sl.stop - sl.start)
    stats = cb_train.average_over_ranks({"a.bn": np.full((2, 3), float(rank + 1), np.float32)})
    q.put((rank, (sl.start, sl.stop), (flat * scale).numpy().copy(), stats["a.bn"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_uneven_shards_give_the_global_mean_gradient_and_averaged_bn_statistics():
    """shard_batch hands the remainder to the first ranks; the global-mean gradient then weights every rank's mean gradient
    by its share of the samples.  Also the BN moving statistics a checkpoint stores are the mean over the ranks."""
    from oracle import ref_train
    w, images, z, pi = _data()
    n, keys = N - 1, sorted(w)
    parts = [ref_train.forward_backward(w, images[:n][s], z[:n][s], pi[:n][s])[1] for s in (slice(0, 4), slice(4, 7))]
    expected = (4 * _flat(parts[0], keys) + 3 * _flat(parts[1], keys)) / 7.0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 90
    procs = [ctx.Process(target=_worker_uneven, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [(0, 4), (4, 7)]
    for _, _, grad, bn in got:
        assert np.allclose(grad, expected, rtol=1e-3, atol=1e-4 * float(np.abs(expected).max()))
        assert np.array_equal(bn, np.full((2, 3), 1.5, np.float32))
    assert np.array_equal(got[0][2], got[1][2])
