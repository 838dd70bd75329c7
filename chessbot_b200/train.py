"""The training step of the reference's train.py (train.py:104-164) over the B200 kernels: `sample_buffer`, `scheduler` and a
`Trainer` whose `train_on_batch` / `fit` do what `model.fit` does there with the model, losses and optimizer of
model.py:19-70 — BatchNormalization in training mode, mean_squared_error + CategoricalCrossentropy(from_logits) + l2(4e-4)
kernel regularizers, SGD(momentum 0.9, nesterov).

PyTorch owns the three flat parameter / gradient / momentum buffers and, with several GPUs, all-reduces the gradient buffer
over NCCL (one process per GPU, data parallel over the batch; BN statistics stay per GPU).  Every computation is a kernel of
libchessbot_b200.so (csrc/train.cu, wgrad.cu, conv.cu) behind include/chessbot_b200.h; there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import CB_NUM_ACTIONS, check, last_error, lib
from .config import Config
from .engine import Engine, _ptr, _stream

PROF_STAGES = ("conv_fwd", "bn_fwd", "heads", "bn_bwd", "wgrad", "dgrad", "optimizer", "input")


def sample_buffer(buffer, batch_size, rng=None):
    """train.py:104-110: uniform sample without replacement -> images, (terminal_values, search_statistics)."""
    rng = np.random.default_rng() if rng is None else rng
    batch = rng.choice(len(buffer), size=batch_size, replace=False)
    images = np.array([buffer[b][0] for b in batch])
    terminal_values = np.array([buffer[b][1][0] for b in batch])
    search_statistics = np.array([buffer[b][1][1] for b in batch])
    return images, (terminal_values, search_statistics)


def scheduler(epoch, lr, config=None):
    """train.py:112-115: learning-rate steps of config.learning_rate."""
    config = Config() if config is None else config
    return config.learning_rate[epoch] if epoch in config.learning_rate else lr


def data_parallel_world():
    """(rank, world) of the data-parallel job: one process per GPU under torch.distributed, else (0, 1)."""
    d = torch.distributed
    return (d.get_rank(), d.get_world_size()) if d.is_available() and d.is_initialized() else (0, 1)


def shard_batch(n, rank=None, world=None):
    """the contiguous slice of a global batch of n samples that rank `rank` trains on (remainder to the first ranks)"""
    r, w = data_parallel_world()
    rank, world = (r if rank is None else rank), (w if world is None else world)
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def allreduce_gradients(flat, local_n=None):
    """Sums the flat gradient buffer over the ranks in place (NCCL over NVLink on the GPUs; gloo in the CPU tests) and returns the
    factor that turns that sum into the gradient of the GLOBAL mean loss.  Every rank holds the mean gradient of its own shard; with
    local_n (samples of this rank's shard) given, shards may be unequal (shard_batch hands the remainder to the first ranks): each
    rank's buffer is weighted by its share local_n / global_n before the sum (one extra scalar all-reduce).  Without local_n the
    shards are taken as equal and the factor is 1 / world."""
    _, world = data_parallel_world()
    if world == 1:
        return 1.0
    if local_n is None:
        torch.distributed.all_reduce(flat)
        return 1.0 / world
    count = torch.tensor([float(local_n)], dtype=torch.float64, device=flat.device)
    torch.distributed.all_reduce(count)
    share = float(local_n) / float(count.item())
    if abs(share * world - 1.0) > 1e-12:
        flat.mul_(share * world)                           # equal shards skip the pass over the buffer
    torch.distributed.all_reduce(flat)
    return 1.0 / world


def average_over_ranks(arrays):
    """element-wise mean over the ranks of a dict of float32 host arrays (BatchNormalization moving statistics before a checkpoint)"""
    _, world = data_parallel_world()
    if world == 1:
        return arrays
    dev = "cuda" if torch.distributed.get_backend() == "nccl" else "cpu"
    out = {}
    for k in sorted(arrays):
        t = torch.as_tensor(np.ascontiguousarray(arrays[k], dtype=np.float32)).to(dev)
        torch.distributed.all_reduce(t)
        out[k] = (t / world).cpu().numpy()
    return out


def tensor_shapes(filters, blocks):
    """name -> Keras-layout shape of every tensor of the model (cb_net_set_tensor / cb_train_set_tensor)."""
    c = filters
    s = {"stem.w": (3, 3, 119, c), "stem.bn": (4, c)}
    for i in range(blocks):
        s[f"res{i}.w1"] = (3, 3, c, c)
        s[f"res{i}.bn1"] = (4, c)
        s[f"res{i}.w2"] = (3, 3, c, c)
        s[f"res{i}.bn2"] = (4, c)
    s.update({"value.conv": (1, 1, c, 1), "value.bn": (4, 1), "value.fc1": (64, c), "value.fc2": (c, 1),
              "policy.conv": (1, 1, c, 2), "policy.bn": (4, 2), "policy.fc": (128, CB_NUM_ACTIONS)})
    return s


class Trainer:
    """One model being trained on one GPU (one per process; `torch.distributed` initialised => data parallel).

    weights: dict of Keras-layout float32 arrays (model.DeviceNetwork / oracle.ref_model.init_weights names);
    batch_size: samples per step on THIS GPU (fixed: the activation buffers of every layer are allocated for it)."""

    def __init__(self, weights, filters, blocks, batch_size, config=None, device=None):
        self.config = Config() if config is None else config
        self.engine = Engine.get(device)
        self.device = self.engine.device
        self.filters, self.blocks, self.batch_size = int(filters), int(blocks), int(batch_size)
        n_reg = C.c_size_t()
        self.n_params = lib.cb_train_param_count(self.filters, self.blocks, C.byref(n_reg))
        self.n_regularized = n_reg.value
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.momentum = torch.zeros_like(self.params)
        self._h = lib.cb_train_create(self.engine._h, self.filters, self.blocks, self.batch_size, 0.3, 1e-3, 0.99,
                                      _ptr(self.params), _ptr(self.grads), _ptr(self.momentum))
        if not self._h:
            raise RuntimeError(f"cb_train_create: {last_error()}")
        self.shapes = tensor_shapes(self.filters, self.blocks)
        self.set_weights(weights)
        self.steps = 0
        self.rank, self.world = data_parallel_world()

    def __del__(self, _destroy=lib.cb_train_destroy):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _destroy(h)

    # ------------------------------------------------------------------ weights
    def set_weights(self, weights, what=0):
        """what=0: the model's tensors; what=2: the optimizer's momentum (same names and layouts)"""
        for name, shape in self.shapes.items():
            a = np.ascontiguousarray(weights[name], dtype=np.float32)
            if a.size != int(np.prod(shape)):
                raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
            check(lib.cb_train_set_tensor(self._h, name.encode(), what, a.ctypes.data, a.size, _stream(self.device.index)), "train_set_tensor")

    def _get(self, what):
        out = {}
        for name, shape in self.shapes.items():
            a = np.empty(shape, dtype=np.float32)
            check(lib.cb_train_get_tensor(self._h, name.encode(), what, a.ctypes.data, a.size, _stream(self.device.index)), "train_get_tensor")
            out[name] = a
        return out

    def get_weights(self):
        """Keras-layout tensors (BN: gamma, beta, moving mean, moving variance): loadable by model.DeviceNetwork for self-play."""
        return self._get(0)

    def get_gradients(self):
        """gradients of the last forward_backward (mean losses, WITHOUT the l2 term; BN rows: dgamma, dbeta, 0, 0)."""
        return self._get(1)

    def get_velocity(self):
        return self._get(2)

    def sync_moving_statistics(self):
        """Data parallel: BatchNormalization's moving mean / variance are per GPU (every rank normalises its own shard, like
        data-parallel Keras); this sets them to the mean over the ranks on every rank, so that a checkpoint does not depend on
        which rank writes it.  Collective: every rank must call it."""
        if self.world == 1:
            return
        w = self.get_weights()
        bn = average_over_ranks({k: a[2:4] for k, a in w.items() if k.endswith("bn") or ".bn" in k})
        for k, rows in bn.items():
            w[k][2:4] = rows
        self.set_weights(w)

    def save(self, path, sync=True):
        """checkpoint for resuming (train.py:124-136 reloads the Keras model with its optimizer state): weights, momentum, step.
        Data parallel: collective (sync=True averages the BN moving statistics over the ranks first); rank 0 writes the file."""
        if sync:
            self.sync_moving_statistics()
        if self.world > 1 and self.rank != 0:
            return
        w, v = self.get_weights(), self.get_velocity()
        np.savez(path, __filters=self.filters, __blocks=self.blocks, __steps=self.steps,
                 **{"w/" + k: a for k, a in w.items()}, **{"v/" + k: a for k, a in v.items()})

    def load(self, path):
        with np.load(path) as f:
            if int(f["__filters"]) != self.filters or int(f["__blocks"]) != self.blocks:
                raise ValueError("checkpoint is for another network shape")
            self.set_weights({k[2:]: f[k] for k in f.files if k.startswith("w/")})
            self.set_weights({k[2:]: f[k] for k in f.files if k.startswith("v/")}, what=2)
            self.steps = int(f["__steps"])

    # ------------------------------------------------------------------ one step
    def forward_backward(self, images, z, pi):
        """images int16 [B,8,8,119], z [B], pi [B,4672] (host arrays or device tensors) -> gradients in self.grads."""
        img = torch.as_tensor(images).to(device=self.device, dtype=torch.int16, non_blocking=True).contiguous()
        zt = torch.as_tensor(z).to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        pt = torch.as_tensor(pi).to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        n = img.shape[0]
        if tuple(img.shape) != (n, 8, 8, 119) or tuple(pt.shape) != (n, CB_NUM_ACTIONS) or zt.numel() != n:
            raise ValueError("expected images [B,8,8,119], terminal values [B], search statistics [B,4672]")
        check(lib.cb_train_forward_backward(self._h, _ptr(img), _ptr(zt), _ptr(pt), n, _stream(self.device.index)), "train_forward_backward")
        self._keep = (img, zt, pt)     # the kernels read them asynchronously
        self._last_n = n

    def apply(self, lr, momentum=None, nesterov=True, l2=None):
        momentum = self.config.momentum if momentum is None else momentum
        l2 = self.config.l2_reg if l2 is None else l2
        # NCCL over NVLink: the only collective of the training step.  Every rank's trainer has a fixed batch, so shards are equal
        # unless the ranks were built with different batch sizes (weighted_shards=True then weights them by their sample counts)
        scale = allreduce_gradients(self.grads, getattr(self, "_last_n", None) if getattr(self, "weighted_shards", False) else None)
        check(lib.cb_train_apply(self._h, float(lr), float(momentum), int(nesterov), float(l2), scale, _stream(self.device.index)), "train_apply")
        self.steps += 1

    def losses(self):
        """{"value": mse, "policy": cross-entropy, "l2": regularization term, "loss": their sum} of the last step (this GPU's batch)."""
        out = (C.c_double * 3)()
        check(lib.cb_train_losses_sync(self._h, out, _stream(self.device.index)), "train_losses")
        return {"value": out[0], "policy": out[1], "l2": out[2], "loss": out[0] + out[1] + out[2]}

    def train_on_batch(self, images, z, pi, lr):
        self.forward_backward(images, z, pi)
        self.apply(lr)
        return self.losses()

    def fit(self, x, y, lr, batch_size=None, shuffle=True, rng=None):
        """model.fit(x, {"value_head": yv, "policy_head": yp}) of train.py:150-164 for one epoch: minibatches of `batch_size`
        (default: the trainer's; Keras' own default would be 32), shuffled.  Returns the mean losses."""
        yv, yp = (y["value_head"], y["policy_head"]) if isinstance(y, dict) else y
        bs = self.batch_size if batch_size is None else batch_size
        if bs != self.batch_size:
            raise ValueError("the trainer was created for another batch size")
        n = len(x)
        order = (np.random.default_rng() if rng is None else rng).permutation(n) if shuffle else np.arange(n)
        logs = []
        for i in range(0, n - bs + 1, bs):
            idx = order[i:i + bs]
            logs.append(self.train_on_batch(np.asarray(x)[idx], np.asarray(yv)[idx], np.asarray(yp)[idx], lr))
        return {k: float(np.mean([l[k] for l in logs])) for k in logs[0]} if logs else {}

    def values(self):
        out = np.empty(self.batch_size, dtype=np.float32)
        check(lib.cb_train_outputs_sync(self._h, out.ctypes.data, _stream(self.device.index)), "train_outputs")
        return out

    # ------------------------------------------------------------------ measurement
    def profile(self, enable):
        check(lib.cb_train_profile(self._h, int(enable)), "train_profile")

    def profile_read(self):
        ms = (C.c_double * len(PROF_STAGES))()
        n = (C.c_int * len(PROF_STAGES))()
        check(lib.cb_train_profile_read_sync(self._h, ms, n), "train_profile_read")
        return {k: (ms[i], n[i]) for i, k in enumerate(PROF_STAGES)}


def train_flops_per_sample(filters, blocks, in_channels=119):
    """algorithmic FLOPs of one sample through forward + data gradient + weight gradient of the 3x3 convolutions."""
    c = filters
    fwd = 2 * 64 * 9 * c * (in_channels + 2 * blocks * c)
    dgrad = 2 * 64 * 9 * c * (2 * blocks * c)           # the stem needs no data gradient
    return fwd + dgrad + fwd
