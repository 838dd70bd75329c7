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
