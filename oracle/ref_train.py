"""ORACLE (test infrastructure). torch-fp32 autograd restatement of ONE training step of the reference:
train.py:150-164 (`model.fit` on a sampled batch) with the model, losses and optimizer of model.py:19-70.

The arithmetic of the reference lives in TensorFlow/Keras (third-party, absent here and un-pinned), so this restates the
*published* Keras semantics: BatchNormalization in training mode normalises with the biased batch
variance, eps 1e-3, and moves its statistics as moving = moving*0.99 + batch*0.01; LeakyReLU slope 0.3; every Conv2D / Dense
kernel carries l2(4e-4) => loss += 4e-4 * sum(w^2); losses "mean_squared_error" (value, tanh output) and
CategoricalCrossentropy(from_logits=True) (policy, target = visit distribution), summed unweighted, each a mean over the batch;
SGD(momentum=0.9, nesterov=True): v = m*v - lr*g; w += m*v - lr*g.
PARITY UNPINNED: the reference ships no weights, golden losses or training tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .ref_model import BN_EPS, SLOPE, num_blocks

BN_MOMENTUM = 0.99
L2 = 4e-4          # config.py:22
MOMENTUM = 0.9     # config.py:25


class _RoundBoth(torch.autograd.Function):
    """value AND incoming gradient rounded to bf16: a tensor the device keeps in bf16 in both passes"""
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _RoundGrad(torch.autograd.Function):
    """identity forward, gradient rounded to bf16 (the residual branch's gradient is stored before it is added)"""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _RoundValue(torch.autograd.Function):
    """value rounded to bf16, gradient passed straight through (bf16 copy of an f32 master weight)"""
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def kernel_names(w):
    """the tensors that carry kernel_regularizer=l2 (model.py:21,24,35,43,47,49,52,56)"""
    return [k for k in w if not k.endswith(("bn", "bn1", "bn2"))]


def forward_backward(w, images, z, pi, device="cpu", l2=L2, precision="fp32", dtype=torch.float32):
    """-> (losses dict, grads dict in the Keras layouts of `w` (BN: [4,C] rows dgamma, dbeta, 0, 0; WITHOUT the l2 term),
    batch statistics dict {bn name: (mean, biased var)}, value [B], logits [B,4672]).

    precision="fp32": the reference's arithmetic.  precision="bf16": the same graph with the tensors the device path keeps in
    bf16 rounded at the same places (tower kernels of the 3x3 convolutions, raw convolution outputs, activations, and in
    the backward pass the gradients of activations / raw outputs / residual branches); everything else stays fp32.  The
    device must agree with THIS variant tightly; its distance to "fp32" is the cost of bf16 storage, not of the kernels."""
    mixed = precision == "bf16"
    rb = _RoundBoth.apply if mixed else (lambda t: t)
    rg = _RoundGrad.apply if mixed else (lambda t: t)
    rv = _RoundValue.apply if mixed else (lambda t: t)
    dev = torch.device(device)
    P = {}
    for k, v in w.items():
        t = torch.tensor(np.asarray(v, dtype=np.float32), device=dev).to(dtype)   # dtype=float64: finite-difference checks only
        if k.endswith(("bn", "bn1", "bn2")):
            P[k] = (t[0].clone().requires_grad_(True), t[1].clone().requires_grad_(True))
        else:
            P[k] = t.requires_grad_(True)
    stats = {}

    def conv2d(x, key, tower=False):
        kern = P[key]
        if kern.dim() != 4:
            kern = kern.reshape(1, 1, kern.shape[0], -1)
        if tower and mixed:
            if key == "stem.w":            # the stem's weights act as bf16 hi + lo pairs (two K passes): ~16 significant bits
                hi = kern.detach().to(torch.bfloat16).to(torch.float32)
                lo = (kern.detach() - hi).to(torch.bfloat16).to(torch.float32)
                kern = kern + (hi + lo - kern).detach()
            else:
                kern = rv(kern)
        y = F.conv2d(x, kern.permute(3, 2, 0, 1), padding=kern.shape[0] // 2)
        return rb(y) if tower else y

    def bn(y, key):
        g, b = P[key]
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        stats[key] = (mean.detach().cpu().numpy(), var.detach().cpu().numpy())
        yh = (y - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
        return g.view(1, -1, 1, 1) * yh + b.view(1, -1, 1, 1)

    x = torch.as_tensor(np.asarray(images, dtype=np.float32)).to(dev).to(dtype).permute(0, 3, 1, 2).contiguous()
    x = rb(F.leaky_relu(bn(conv2d(x, "stem.w", True), "stem.bn"), SLOPE))
    for i in range(num_blocks(w)):
        y = rb(F.leaky_relu(bn(conv2d(x, f"res{i}.w1", True), f"res{i}.bn1"), SLOPE))
        y = bn(conv2d(y, f"res{i}.w2", True), f"res{i}.bn2")
        x = rb(F.leaky_relu(rg(x) + y, SLOPE))
    B = x.shape[0]
    v = F.leaky_relu(bn(conv2d(x, "value.conv"), "value.bn"), SLOPE).permute(0, 2, 3, 1).reshape(B, 64)
    v = F.leaky_relu(v @ P["value.fc1"], SLOPE)
    value = torch.tanh(v @ P["value.fc2"].reshape(-1, 1)).reshape(B)
    p = F.leaky_relu(bn(conv2d(x, "policy.conv"), "policy.bn"), SLOPE).permute(0, 2, 3, 1).reshape(B, 128)
    logits = p @ P["policy.fc"]
    zt = torch.as_tensor(np.asarray(z, dtype=np.float32)).to(dev).to(dtype)
    pt = torch.as_tensor(np.asarray(pi, dtype=np.float32)).to(dev).to(dtype)
    loss_v = ((value - zt) ** 2).mean()
    loss_p = -(pt * F.log_softmax(logits, dim=1)).sum(dim=1).mean()
    reg = sum((P[k] ** 2).sum() for k in kernel_names(w)) * l2
    (loss_v + loss_p).backward()
    grads = {}
    for k, v_ in w.items():
        if k.endswith(("bn", "bn1", "bn2")):
            g, b = P[k]
            c = g.shape[0]
            grads[k] = np.stack([g.grad.cpu().numpy(), b.grad.cpu().numpy(), np.zeros(c, np.float32), np.zeros(c, np.float32)]).astype(np.float32)
        else:
            grads[k] = P[k].grad.cpu().numpy().reshape(np.asarray(v_).shape)
    lv, lp, lr_ = loss_v.item(), loss_p.item(), reg.item()
    losses = {"value": lv, "policy": lp, "l2": lr_, "loss": lv + lp + lr_}
    return losses, grads, stats, value.detach().cpu().numpy(), logits.detach().cpu().numpy()


def sgd_step(w, grads, velocity, stats, lr, momentum=MOMENTUM, nesterov=True, l2=L2, bn_momentum=BN_MOMENTUM):
    """Keras SGD on every trainable tensor (+ the l2 gradient 2*l2*w on kernels) and the BN moving-average update.
    velocity: dict like `w` (zeros at the start).  Returns (new weights, new velocity)."""
    new_w, new_v = {}, {}
    for k, val in w.items():
        val = np.asarray(val, dtype=np.float32)
        g = np.asarray(grads[k], dtype=np.float32)
        vel = np.asarray(velocity[k], dtype=np.float32) if velocity is not None else np.zeros_like(val)
        if k.endswith(("bn", "bn1", "bn2")):
            out, ov = val.copy(), vel.copy()
            for r in range(2):
                vr = momentum * vel[r] - lr * g[r]
                ov[r] = vr
                out[r] = val[r] + (momentum * vr - lr * g[r] if nesterov else vr)
            mean, var = stats[k]
            out[2] = val[2] * bn_momentum + mean * (1 - bn_momentum)
            out[3] = val[3] * bn_momentum + var * (1 - bn_momentum)
            new_w[k], new_v[k] = out.astype(np.float32), ov.astype(np.float32)
        else:
            g = g + 2 * l2 * val
            vr = momentum * vel - lr * g
            new_v[k] = vr.astype(np.float32)
            new_w[k] = (val + (momentum * vr - lr * g if nesterov else vr)).astype(np.float32)
    return new_w, new_v


def train_step(w, velocity, images, z, pi, lr, device="cpu", precision="fp32", **kw):
    """One step of model.fit on one batch: -> (losses, new weights, new velocity, grads)."""
    losses, grads, stats, _, _ = forward_backward(w, images, z, pi, device=device, l2=kw.get("l2", L2), precision=precision)
    new_w, new_v = sgd_step(w, grads, velocity, stats, lr, **kw)
    return losses, new_w, new_v, grads


def synthetic_targets(n, seed, num_actions=4672, support=30):
    """z in {-1, 0, 1} and sparse visit distributions (train.py:13-19 shapes) for synthetic batches."""
    rng = np.random.default_rng(seed)
    z = rng.integers(-1, 2, n).astype(np.float32)
    pi = np.zeros((n, num_actions), dtype=np.float32)
    for i in range(n):
        idx = rng.choice(num_actions, size=support, replace=False)
        cnt = rng.integers(1, 50, support).astype(np.float64)
        pi[i, idx] = (cnt / cnt.sum()).astype(np.float32)
    return z, pi
