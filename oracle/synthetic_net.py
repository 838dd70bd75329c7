"""ORACLE (test infrastructure). NumPy statement of the *synthetic network*.

The reference's tests drive the search with a mock `network(image) -> {"value_head",
"policy_head"}` closure (tests/mcts_v2.py:51-67; main.ipynb:1311-1315).  The device build has the
same fake backend as a kernel (`cb_synth_forward`): outputs are an integer hash of the input planes,
so that tree parity can be checked bit for bit without the convolution tower in the loop.

  h      = sum_i plane[i] * ((i * 2654435761 + 0x9E3779B9) mod 2^32)                 (mod 2^32),
           i = (row*8 + col)*119 + channel  (NHWC order, 119 real channels)
  value  = (((mix(h ^ seed) >> 8) & 0xFFFF) - 32768) / 32768                        (exact f32)
  logit_a= (((mix(h ^ seed ^ ((a+1) * 0x9E3779B1)) >> 16) & 0xFF) - 128) / 32       (exact bf16)
  mix    = murmur3 fmix32
"""
import numpy as np

_COEF = ((np.arange(8 * 8 * 119, dtype=np.uint64) * 2654435761 + 0x9E3779B9) & 0xFFFFFFFF).astype(np.uint64)
_A = ((np.arange(1, 4673, dtype=np.uint64) * 0x9E3779B1) & 0xFFFFFFFF).astype(np.uint32)


def fmix32(x):
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def plane_hash(image):
    v = np.asarray(image).reshape(-1).astype(np.int64).astype(np.uint64)
    return np.uint32(int((v * _COEF).sum() & 0xFFFFFFFF))


def outputs(image, seed=0):
    h = np.uint32(plane_hash(image) ^ np.uint32(seed))
    value = np.float32((int((fmix32(h) >> 8) & 0xFFFF) - 32768) / 32768.0)
    logits = ((((fmix32(np.uint32(h) ^ _A) >> 16) & 0xFF).astype(np.int32) - 128) / 32.0).astype(np.float32)
    return value, logits


def as_network(seed=0):
    def net(image):
        img = np.asarray(image)
        vs, ps = zip(*(outputs(img[i], seed) for i in range(img.shape[0])))
        return {"value_head": np.asarray(vs, dtype=np.float32).reshape(-1, 1),
                "policy_head": np.stack(ps).astype(np.float32)}
    return net
