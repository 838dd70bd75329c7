"""CPU tests of the training-step oracle (oracle/ref_train.py) and of the host side of chessbot_b200/train.py.

The reference ships no training tests, golden losses or weights (parity unpinned), so the oracle is pinned
against first principles instead: finite differences of its own loss, hand-computed known answers for the two losses,
the l2 term and the Keras SGD(momentum, nesterov) update, and the BatchNormalization moving-average rule."""
import numpy as np
import pytest
import torch

from oracle import ref_model, ref_train


def _tiny(seed=0, n=6, filters=8, blocks=1):
    w = ref_model.init_weights(filters, blocks, seed=seed, random_bn=True)
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 2, (n, 8, 8, 119)).astype(np.int16)
    images[..., 113] = rng.integers(0, 120, (n, 1, 1))
    images[..., 118] = rng.integers(0, 40, (n, 1, 1))
    z, pi = ref_train.synthetic_targets(n, seed)
    return w, images, z, pi


def test_gradients_match_finite_differences_of_the_loss():
    w, images, z, pi = _tiny(1)
    losses, g, _, _, _ = ref_train.forward_backward(w, images, z, pi)
    for key, idx in (("res0.w1", (1, 1, 3, 4)), ("res0.w2", (0, 2, 1, 1)), ("stem.bn", (0, 5)), ("res0.bn2", (1, 2)),
                     ("policy.fc", (10, int(pi[0].argmax()))), ("value.fc1", (3, 2)), ("value.fc2", (4, 0)), ("policy.conv", (0, 0, 3, 1))):
        eps = 1e-4
        vals = []
        for sgn in (1, -1):
            w2 = {k: np.array(v, copy=True) for k, v in w.items()}
            w2[key][idx] += sgn * eps
            l2 = ref_train.forward_backward(w2, images, z, pi, dtype=torch.float64)[0]
            vals.append(l2["value"] + l2["policy"])
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - g[key][idx]) <= 2e-3 * abs(g[key][idx]) + 2e-5, (key, fd, g[key][idx])


def test_losses_known_answers():
    """mean_squared_error, CategoricalCrossentropy(from_logits) and the l2 term, recomputed by hand from the forward outputs."""
    w, images, z, pi = _tiny(2)
    losses, _, _, value, logits = ref_train.forward_backward(w, images, z, pi)
    assert losses["value"] == pytest.approx(float(np.mean((value - z) ** 2)), rel=1e-5)
    lse = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1)
    assert losses["policy"] == pytest.approx(float(np.mean(-(pi * (logits - lse[:, None])).sum(1))), rel=1e-5)
    reg = 4e-4 * sum(float((np.asarray(v, np.float64) ** 2).sum()) for k, v in w.items() if "bn" not in k)
    assert losses["l2"] == pytest.approx(reg, rel=1e-5)
    assert losses["loss"] == pytest.approx(losses["value"] + losses["policy"] + losses["l2"], rel=1e-9)


def test_training_mode_batchnorm_uses_batch_statistics_not_moving_ones():
    w, images, z, pi = _tiny(3)
    w2 = {k: np.array(v, copy=True) for k, v in w.items()}
    for k in w2:
        if "bn" in k:
            w2[k][2] += 5.0      # moving mean / variance must not matter in training mode
            w2[k][3] *= 7.0
    a = ref_train.forward_backward(w, images, z, pi)[0]
    b = ref_train.forward_backward(w2, images, z, pi)[0]
    assert a["value"] == b["value"] and a["policy"] == b["policy"]


def test_sgd_nesterov_and_moving_average_known_answers():
    """Keras SGD(momentum=0.9, nesterov=True): v = m v - lr g; w += m v - lr g, with g = grad + 2 l2 w on kernels only."""
    w = {"value.fc2": np.array([[1.0], [-2.0]], np.float32), "value.bn": np.array([[1.0], [0.5], [0.2], [2.0]], np.float32)}
    g = {"value.fc2": np.array([[0.5], [0.25]], np.float32), "value.bn": np.array([[0.1], [-0.2], [0], [0]], np.float32)}
    vel = {"value.fc2": np.array([[0.1], [0.0]], np.float32), "value.bn": np.zeros((4, 1), np.float32)}
    stats = {"value.bn": (np.array([1.2], np.float32), np.array([4.0], np.float32))}
    nw, nv = ref_train.sgd_step(w, g, vel, stats, lr=0.1, momentum=0.9, nesterov=True, l2=0.5)
    gk = g["value.fc2"] + 2 * 0.5 * w["value.fc2"]                       # [[1.5], [-1.75]]
    v_new = 0.9 * vel["value.fc2"] - 0.1 * gk                            # [[-0.06], [0.175]]
    assert np.allclose(nv["value.fc2"], v_new)
    assert np.allclose(nw["value.fc2"], w["value.fc2"] + 0.9 * v_new - 0.1 * gk)
    assert np.allclose(nw["value.bn"][:2, 0], [1.0 - 0.1 * 0.1 * 1.9, 0.5 + 0.1 * 0.2 * 1.9])     # no l2 on gamma / beta
    assert np.allclose(nw["value.bn"][2:, 0], [0.2 * 0.99 + 1.2 * 0.01, 2.0 * 0.99 + 4.0 * 0.01])
    plain, _ = ref_train.sgd_step(w, g, vel, stats, lr=0.1, momentum=0.9, nesterov=False, l2=0.5)
    assert np.allclose(plain["value.fc2"], w["value.fc2"] + v_new)


def test_bf16_rounding_variant_stays_close_to_fp32():
    w, images, z, pi = _tiny(4, n=8, filters=16)
    l32, g32, *_ = ref_train.forward_backward(w, images, z, pi)
    l16, g16, *_ = ref_train.forward_backward(w, images, z, pi, precision="bf16")
    assert abs(l32["policy"] - l16["policy"]) < 1e-2 and abs(l32["value"] - l16["value"]) < 2e-2
    for k in ("stem.w", "res0.w1", "policy.fc"):
        a, b = g16[k].ravel().astype(np.float64), g32[k].ravel().astype(np.float64)
        assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b)) > 0.98, k
    # the straight-through rounding must not change the gradient's scale (stem hi + lo weights included)
    a, b = g16["stem.w"].ravel().astype(np.float64), g32["stem.w"].ravel().astype(np.float64)
    assert abs(a @ b / (b @ b) - 1) < 0.05


def test_host_side_sample_buffer_scheduler_and_layout():
    """train.py:104-115 host logic and the flat parameter layout of the C ABI (no GPU needed)."""
    import ctypes as C

    from chessbot_b200 import train as cb_train
    from chessbot_b200._lib import lib
    rng = np.random.default_rng(0)
    buffer = [(np.full((8, 8, 119), i, np.int16), (float(i % 3 - 1), np.eye(4672)[i])) for i in range(10)]
    x, (yv, yp) = cb_train.sample_buffer(buffer, 4, rng)
    assert x.shape == (4, 8, 8, 119) and yv.shape == (4,) and yp.shape == (4, 4672)
    ids = x[:, 0, 0, 0]
    assert len(set(ids.tolist())) == 4 and all(yp[j, ids[j]] == 1 for j in range(4)) and all(yv[j] == ids[j] % 3 - 1 for j in range(4))
    assert cb_train.scheduler(0, 5.0) == 0.1 and cb_train.scheduler(14000, 0.1) == 0.01 and cb_train.scheduler(17, 0.3) == 0.3
    for filters, blocks in ((48, 4), (128, 6), (256, 20)):
        shapes = cb_train.tensor_shapes(filters, blocks)
        n_reg = C.c_size_t()
        n = lib.cb_train_param_count(filters, blocks, C.byref(n_reg))
        trainable = sum(int(np.prod(s)) if "bn" not in k else 2 * s[1] for k, s in shapes.items())
        kernels = sum(int(np.prod(s)) for k, s in shapes.items() if "bn" not in k)
        assert (n, n_reg.value) == (trainable, kernels)
        assert set(shapes) == set(ref_model.init_weights(filters, blocks))
    assert cb_train.train_flops_per_sample(256, 20) == 2 * (2 * 64 * 9 * 256 * (119 + 40 * 256)) + 2 * 64 * 9 * 256 * 40 * 256
