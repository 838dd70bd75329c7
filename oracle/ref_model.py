"""ORACLE (test infrastructure). torch-fp32 restatement of the forward graph of model.py:19-56.

The arithmetic of the reference lives in TensorFlow/Keras/TF-TRT (third-party, absent here and
un-pinned), so this restates the *published* Keras layer semantics:
Conv2D(padding="same", use_bias=False) with HWIO kernels on NHWC data; BatchNormalization()
inference y = gamma*(x-mean)/sqrt(var+1e-3)+beta; LeakyReLU() slope 0.3; Flatten() of NHWC =>
index (h*8+w)*C+c; Dense(use_bias=False) kernels [in,out]; tanh on the value output.
PARITY UNPINNED: the reference ships no weights, golden outputs or model tests.

Weights dict (all float32 numpy, the layout the C-ABI `cb_net_load` takes):
  stem.w [3,3,Cin,C]; stem.bn [4,C] (gamma,beta,mean,var)
  res{i}.w1 / res{i}.w2 [3,3,C,C]; res{i}.bn1 / res{i}.bn2 [4,C]
  value.conv [C,1]... stored as [1,1,C,1]; value.bn [4,1]; value.fc1 [64,C]; value.fc2 [C,1]
  policy.conv [1,1,C,2]; policy.bn [4,2]; policy.fc [128,4672]
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
SLOPE = 0.3
NUM_ACTIONS = 4672


def glorot_uniform(gen, shape, fan_in, fan_out):
    lim = float(np.sqrt(6.0 / (fan_in + fan_out)))
    return ((torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim).numpy()


def init_weights(filters=48, blocks=4, in_channels=119, seed=0, random_bn=False):
    """Keras random-init state (glorot_uniform kernels, BN gamma=1 beta=0 mean=0 var=1).
    random_bn=True perturbs the BN statistics so that the fused scale/shift path is exercised."""
    g = torch.Generator().manual_seed(seed)
    C = filters

    def bn(c):
        if random_bn:
            gamma = 1 + 0.2 * torch.randn(c, generator=g)
            beta = 0.1 * torch.randn(c, generator=g)
            mean = 0.1 * torch.randn(c, generator=g)
            var = 0.5 + torch.rand(c, generator=g)
            return torch.stack([gamma, beta, mean, var]).numpy().astype(np.float32)
        return np.stack([np.ones(c), np.zeros(c), np.zeros(c), np.ones(c)]).astype(np.float32)

    w = {"stem.w": glorot_uniform(g, (3, 3, in_channels, C), 9 * in_channels, 9 * C), "stem.bn": bn(C)}
    for i in range(blocks):
        w[f"res{i}.w1"] = glorot_uniform(g, (3, 3, C, C), 9 * C, 9 * C)
        w[f"res{i}.bn1"] = bn(C)
        w[f"res{i}.w2"] = glorot_uniform(g, (3, 3, C, C), 9 * C, 9 * C)
        w[f"res{i}.bn2"] = bn(C)
    w["value.conv"] = glorot_uniform(g, (1, 1, C, 1), C, 1)
    w["value.bn"] = bn(1)
    w["value.fc1"] = glorot_uniform(g, (64, C), 64, C)
    w["value.fc2"] = glorot_uniform(g, (C, 1), C, 1)
    w["policy.conv"] = glorot_uniform(g, (1, 1, C, 2), C, 2)
    w["policy.bn"] = bn(2)
    w["policy.fc"] = glorot_uniform(g, (128, NUM_ACTIONS), 128, NUM_ACTIONS)
    return w


def num_blocks(w):
    return sum(1 for k in w if k.endswith(".w1"))


def _conv(x, w_hwio, device):
    k = torch.from_numpy(np.ascontiguousarray(w_hwio)).to(device).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, k, padding=k.shape[-1] // 2)


def _bn(x, bn, device):
    g, b, m, v = (torch.from_numpy(np.ascontiguousarray(bn[i])).to(device).view(1, -1, 1, 1) for i in range(4))
    return g * (x - m) / torch.sqrt(v + BN_EPS) + b


def forward(w, images, device="cpu", return_tower=False):
    """images: [B,8,8,Cin] (any numeric dtype) -> (value f32 [B], logits f32 [B,4672])."""
    with torch.no_grad():
        x = torch.as_tensor(np.asarray(images, dtype=np.float32)).to(device).permute(0, 3, 1, 2).contiguous()
        x = F.leaky_relu(_bn(_conv(x, w["stem.w"], device), w["stem.bn"], device), SLOPE)
        for i in range(num_blocks(w)):
            y = F.leaky_relu(_bn(_conv(x, w[f"res{i}.w1"], device), w[f"res{i}.bn1"], device), SLOPE)
            y = _bn(_conv(y, w[f"res{i}.w2"], device), w[f"res{i}.bn2"], device)
            x = F.leaky_relu(x + y, SLOPE)
        B = x.shape[0]
        v = F.leaky_relu(_bn(_conv(x, w["value.conv"], device), w["value.bn"], device), SLOPE)
        v = v.permute(0, 2, 3, 1).reshape(B, 64)
        v = F.leaky_relu(v @ torch.from_numpy(w["value.fc1"]).to(device), SLOPE)
        v = torch.tanh(v @ torch.from_numpy(w["value.fc2"]).to(device)).reshape(B)
        p = F.leaky_relu(_bn(_conv(x, w["policy.conv"], device), w["policy.bn"], device), SLOPE)
        p = p.permute(0, 2, 3, 1).reshape(B, 128)
        logits = p @ torch.from_numpy(w["policy.fc"]).to(device)
        if return_tower:
            return v.cpu().numpy(), logits.cpu().numpy(), x.permute(0, 2, 3, 1).contiguous().cpu().numpy()
        return v.cpu().numpy(), logits.cpu().numpy()


def forward_bf16_rounding(w, images, device="cpu", stream_hi_lo=True):
    """The same graph with the roundings of the device datapath applied where the device applies them, everything else in fp32: what an
    ideal bf16-operand tensor-core tower computes.  Convolution operands (activations as stored in HBM, packed weights) are bf16, products
    accumulate in fp32, BN / skip / LeakyReLU run in fp32 on the accumulator; the residual stream is kept as bf16 hi + lo (~fp32) between
    blocks (stream_hi_lo=False: rounded to bf16, the round-1 datapath); the stem sees exact integers with bf16 hi + lo weights (~fp32); the
    fused 1x1 head convolutions read the fp32 epilogue value.  The distance device <-> this oracle isolates kernel defects (it should be
    accumulation-order noise); the distance of this oracle <-> `forward` is what bf16 operands cost, whatever the kernel."""
    def rb(t):
        return t.to(torch.bfloat16).to(torch.float32)

    def fold(bn):
        g, b, m, v = (torch.from_numpy(np.ascontiguousarray(bn[i])).to(device).double() for i in range(4))
        s = g / torch.sqrt(v + BN_EPS)
        return s.float().view(1, -1, 1, 1), (b - m * s).float().view(1, -1, 1, 1)

    def kern(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device).permute(3, 2, 0, 1).contiguous()

    with torch.no_grad():
        x = torch.as_tensor(np.asarray(images, dtype=np.float32)).to(device).permute(0, 3, 1, 2).contiguous()
        s, b = fold(w["stem.bn"])
        stream = F.leaky_relu(F.conv2d(x, kern(w["stem.w"]), padding=1) * s + b, SLOPE)
        for i in range(num_blocks(w)):
            s1, b1 = fold(w[f"res{i}.bn1"])
            y = rb(F.leaky_relu(F.conv2d(rb(stream), rb(kern(w[f"res{i}.w1"])), padding=1) * s1 + b1, SLOPE))
            s2, b2 = fold(w[f"res{i}.bn2"])
            y = F.conv2d(y, rb(kern(w[f"res{i}.w2"])), padding=1) * s2 + b2
            skip = stream if stream_hi_lo else rb(stream)
            if stream_hi_lo:
                hi = rb(stream)
                skip = hi + rb(stream - hi)
            stream = F.leaky_relu(skip + y, SLOPE)
        x = stream
        B = x.shape[0]
        sv, bv = fold(w["value.bn"])
        v = F.leaky_relu(F.conv2d(x, kern(w["value.conv"])) * sv + bv, SLOPE).permute(0, 2, 3, 1).reshape(B, 64)
        v = F.leaky_relu(v @ torch.from_numpy(w["value.fc1"]).to(device), SLOPE)
        v = torch.tanh(v @ torch.from_numpy(w["value.fc2"]).to(device)).reshape(B)
        sp, bp = fold(w["policy.bn"])
        p = F.leaky_relu(F.conv2d(x, kern(w["policy.conv"])) * sp + bp, SLOPE).permute(0, 2, 3, 1).reshape(B, 128)
        return v.cpu().numpy(), (p @ torch.from_numpy(w["policy.fc"]).to(device)).cpu().numpy()


def calibrate_bn(w, images, device="cpu"):
    """Returns a copy of `w` whose BatchNorm moving statistics are the batch statistics of `images`,
    layer by layer (what training converges to): activations and logits become O(1), the regime the
    2e-2 absolute tolerance of BASELINE.json's north_star is stated for.  Keras random-init BN
    (mean 0, var 1) leaves the raw integer planes un-normalised and logits in the tens."""
    w = {k: np.array(v, copy=True) for k, v in w.items()}

    def fit(y, key):
        c = y.shape[1]
        m = y.mean(dim=(0, 2, 3)).cpu().numpy()
        v = y.var(dim=(0, 2, 3), unbiased=False).cpu().numpy()
        w[key] = np.stack([np.ones(c), np.zeros(c), m, np.maximum(v, 1e-6)]).astype(np.float32)

    with torch.no_grad():
        x = torch.as_tensor(np.asarray(images, dtype=np.float32)).to(device).permute(0, 3, 1, 2).contiguous()
        y = _conv(x, w["stem.w"], device)
        fit(y, "stem.bn")
        x = F.leaky_relu(_bn(y, w["stem.bn"], device), SLOPE)
        for i in range(num_blocks(w)):
            y = _conv(x, w[f"res{i}.w1"], device)
            fit(y, f"res{i}.bn1")
            y = F.leaky_relu(_bn(y, w[f"res{i}.bn1"], device), SLOPE)
            y = _conv(y, w[f"res{i}.w2"], device)
            fit(y, f"res{i}.bn2")
            x = F.leaky_relu(x + _bn(y, w[f"res{i}.bn2"], device), SLOPE)
        fit(_conv(x, w["value.conv"], device), "value.bn")
        fit(_conv(x, w["policy.conv"], device), "policy.bn")
    return w


def as_network(w, device="cpu"):
    """The reference's `network` contract (model.py:77-79): callable(f32[1,8,8,119]) -> dict."""
    def net(image):
        v, p = forward(w, image, device)
        return {"value_head": v.reshape(-1, 1), "policy_head": p}
    return net


def flops_per_position(filters, blocks, in_channels=119):
    C = filters
    return 2 * 64 * 9 * C * (in_channels + 2 * blocks * C) + 2 * 64 * C * 3 + 2 * 64 * C + 2 * C + 2 * 128 * NUM_ACTIONS
