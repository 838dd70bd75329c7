#!/usr/bin/env python
"""Error side of the "path past the dense-bf16 roofline" study: the tower of model.py:19-56 at the
benched shape (20x256, bench.trained_like_bn weights, seeded random-playout positions) evaluated in torch with the
roundings of each candidate datapath emulated at the places the device would apply them, against the torch-fp32 oracle.

    python scripts/precision_variants.py [--n 256] [--net 20x256] [--device cpu|cuda] [--out profiles/r2_precision_variants.json]

Variants (what the tensor cores would see / what HBM would hold):
  bf16            the shipped tower: conv operands bf16 (activations stored bf16, weights bf16), f32 accumulate, f32 epilogue,
                  residual stream stored bf16
  bf16+f32skip    same MMAs; the residual stream also kept in f32 (bf16 hi + bf16 lo in HBM) for the skip add
  fp8a            + first conv of every block on e4m3 operands (per-output-channel weight scales, per-tensor activation scale)
  fp8ab           + both convs of every block on e4m3
  wino_bf16       Winograd F(2x2,3x3): f32 input / weight / output transforms, the 16 element-wise GEMMs on bf16 operands
  wino_bf16+f32skip
Test / study infrastructure: uses oracle/ (never imported by the product).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SLOPE, EPS = 0.3, 1e-3


def rb(x):
    return x.to(torch.bfloat16).to(torch.float32)


def r8(x, scale):
    """e4m3 (max 448) with a scale: value = fp8(x / scale) * scale"""
    return (x / scale).clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float32) * scale


def bn_fold(bn, dev):
    g, b, m, v = (torch.from_numpy(np.ascontiguousarray(bn[i])).to(dev).double() for i in range(4))
    s = g / torch.sqrt(v + EPS)
    return s.float().view(1, -1, 1, 1), (b - m * s).float().view(1, -1, 1, 1)


def kernel(w, dev):
    return torch.from_numpy(np.ascontiguousarray(w)).to(dev).permute(3, 2, 0, 1).contiguous()


# Winograd F(2x2, 3x3) matrices (Lavin & Gray)
_BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
_G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float32)
_AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)


def conv_winograd(x, k, round_op):
    """x [B,C,8,8] f32, k [Co,Ci,3,3] f32 -> 'same' convolution through F(2x2,3x3) with the GEMM operands rounded by round_op"""
    B, C, _, _ = x.shape
    dev = x.device
    BT, G, AT = _BT.to(dev), _G.to(dev), _AT.to(dev)
    xp = F.pad(x, (1, 1, 1, 1))
    d = xp.unfold(2, 4, 2).unfold(3, 4, 2)                      # [B,C,4,4,4,4] tiles (ty,tx) x (4x4)
    V = torch.einsum("ij,bcyxjk,lk->bcyxil", BT, d, BT)         # B^T d B
    U = torch.einsum("ij,ocjk,lk->ocil", G, k, G)               # G g G^T
    V, U = round_op(V), round_op(U)
    M = torch.einsum("bcyxil,ocil->boyxil", V, U)
    Y = torch.einsum("ij,boyxjk,lk->boyxil", AT, M, AT)         # [B,Co,4,4,2,2]
    return Y.permute(0, 1, 2, 4, 3, 5).reshape(B, -1, 8, 8)


def tower(w, images, variant, dev, blocks):
    f32skip = "f32skip" in variant
    wino = variant.startswith("wino")
    fp8_first = variant.startswith("fp8a")
    fp8_both = variant.startswith("fp8ab")
    exact = variant == "fp32"

    def conv(x, key, fp8=False):
        k = kernel(w[key], dev)
        if exact:
            return F.conv2d(x, k, padding=1)
        if wino:
            return conv_winograd(x, k, rb)
        if fp8:
            ks = k.abs().amax(dim=(1, 2, 3), keepdim=True) / 448.0          # per output channel: folds into the BN scale
            xs = x.abs().max() / 448.0                                      # per tensor (a trained network would calibrate it)
            return F.conv2d(r8(x, xs), r8(k, ks), padding=1)
        return F.conv2d(rb(x), rb(k), padding=1)

    with torch.no_grad():
        x = torch.as_tensor(np.asarray(images, dtype=np.float32)).to(dev).permute(0, 3, 1, 2).contiguous()
        s, b = bn_fold(w["stem.bn"], dev)
        # stem: integer planes exact on the device (bf16 hi + remainder channels), weights bf16 hi + lo: f32-like
        x = F.leaky_relu(F.conv2d(x, kernel(w["stem.w"], dev), padding=1) * s + b, SLOPE)
        stream = x                                  # f32 value of the residual stream as the epilogue computed it
        for i in range(blocks):
            xin = stream if exact else rb(stream)   # what HBM holds for the next conv's operand
            s1, b1 = bn_fold(w[f"res{i}.bn1"], dev)
            y = F.leaky_relu(conv(xin, f"res{i}.w1", fp8_first or fp8_both) * s1 + b1, SLOPE)
            if not exact:
                y = rb(y)
            s2, b2 = bn_fold(w[f"res{i}.bn2"], dev)
            y = conv(y, f"res{i}.w2", fp8_both) * s2 + b2
            skip = stream if (exact or f32skip) else rb(stream)
            stream = F.leaky_relu(skip + y, SLOPE)
        x = stream                                  # the fused 1x1 head convolutions read the f32 epilogue value
        B = x.shape[0]
        sv, bv = bn_fold(w["value.bn"], dev)
        v = F.leaky_relu(F.conv2d(x, kernel(w["value.conv"], dev)) * sv + bv, SLOPE).permute(0, 2, 3, 1).reshape(B, 64)
        v = F.leaky_relu(v @ torch.from_numpy(w["value.fc1"]).to(dev), SLOPE)
        v = torch.tanh(v @ torch.from_numpy(w["value.fc2"]).to(dev)).reshape(B)
        sp, bp = bn_fold(w["policy.bn"], dev)
        p = F.leaky_relu(F.conv2d(x, kernel(w["policy.conv"], dev)) * sp + bp, SLOPE).permute(0, 2, 3, 1).reshape(B, 128)
        return v.cpu().numpy(), (p @ torch.from_numpy(w["policy.fc"]).to(dev)).cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--net", default="20x256")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--variants", default="bf16,bf16+f32skip,fp8a+f32skip,fp8ab+f32skip,wino_bf16,wino_bf16+f32skip")
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import bench
    from chessbot_b200.model import keras_init_weights
    from oracle import ref_gameimage
    from gpu_util import oracle_board
    filters, blocks = bench.parse_net(a.net)
    w = bench.trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks)
    boards = bench.random_positions_device(a.n, seed=a.seed)
    images = np.stack([ref_gameimage.board_to_image(oracle_board("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
                                                                   " ".join(m.uci() for m in b.move_stack)), 8) for b in boards])
    t0 = time.time()
    v0, p0 = tower(w, images, "fp32", a.device, blocks)
    print(f"fp32 oracle: {time.time() - t0:.1f} s, |logit| max {np.abs(p0).max():.3f}, std {p0.std():.3f}", flush=True)
    rows = []
    for var in a.variants.split(","):
        t0 = time.time()
        v, p = tower(w, images, var, a.device, blocks)
        dp, dv = np.abs(p - p0), np.abs(v - v0)
        row = {"variant": var, "n": a.n, "net": a.net, "max_dlogit": float(dp.max()), "p999_dlogit": float(np.quantile(dp, 0.999)),
               "rms_dlogit": float(np.sqrt(np.mean(dp ** 2))), "max_dvalue": float(dv.max()), "top1_agree": float((p.argmax(1) == p0.argmax(1)).mean()),
               "holds_2e-2": bool(dp.max() <= 2e-2 and dv.max() <= 2e-2)}
        rows.append(row)
        print(json.dumps(row), f"({time.time() - t0:.1f} s)", flush=True)
    if a.out:
        json.dump({"what": __doc__.split("\n\n")[0], "logit_absmax_fp32": float(np.abs(p0).max()), "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
