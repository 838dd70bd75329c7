"""Training step parity (-m gpu): SURVEY.md §8(f) row 1, train.py:150-164 / model.py:59-70 through the C ABI.

1. the tcgen05 weight-gradient kernel against the plain CUDA-core kernel on the same bf16 inputs and against torch;
2. one whole training step (loss, gradients, updated weights) against the torch-fp32 autograd oracle (oracle/ref_train.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from chessbot_b200.engine import Engine
    return Engine.get()


@pytest.mark.parametrize("cin_pad,n,boards", [(64, 64, 6), (128, 128, 7), (128, 256, 4), (256, 256, 10), (256, 256, 700), (128, 64, 3)])
def test_wgrad_against_cuda_core_kernel_and_torch(engine, cin_pad, n, boards):
    from chessbot_b200.engine import nhwc_to_act
    g = torch.Generator(device="cuda").manual_seed(boards + n)
    x = torch.randn(boards, 8, 8, cin_pad, device="cuda", generator=g).to(torch.bfloat16)
    dy = (torch.randn(boards, 8, 8, n, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    a_in, a_dy = nhwc_to_act(x), nhwc_to_act(dy)
    out_t = engine.debug_wgrad(a_in, a_dy, boards, cin_pad, n, use_reference=False)
    out_r = engine.debug_wgrad(a_in, a_dy, boards, cin_pad, n, use_reference=True)
    torch.cuda.synchronize()
    assert torch.isfinite(out_t).all()
    scale = float(out_r.abs().max())
    assert scale > 1e-2
    assert float((out_t - out_r).abs().max()) <= 2e-4 * scale + 1e-4, (float((out_t - out_r).abs().max()), scale)
    # torch: grad of conv2d w.r.t. its weight = the same sum (cross-correlation, "same" padding)
    torch.backends.cudnn.allow_tf32 = False
    xt = x.float().permute(0, 3, 1, 2).contiguous()
    w = torch.zeros(n, cin_pad, 3, 3, device="cuda", requires_grad=True)
    y = torch.nn.functional.conv2d(xt, w, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2).contiguous())
    ref = w.grad.permute(2, 3, 1, 0).reshape(9, cin_pad, n)        # [kh*3+kw][ci][co]
    assert float((out_t - ref).abs().max()) <= 1e-3 * scale + 1e-3


def _batch(n, seed):
    from gpu_util import oracle_board, random_games
    from oracle import ref_gameimage, ref_train
    games = random_games(n, seed=seed, max_plies=100)
    images = np.stack([ref_gameimage.board_to_image(oracle_board(f, m), 8) for f, m in games]).astype(np.int16)
    z, pi = ref_train.synthetic_targets(n, seed)
    return images, z, pi


def _rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


def _proj(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (b @ b + 1e-30))


# Two oracles (oracle/ref_train.py).  precision="fp32" is the reference's own arithmetic.  A weight gradient is a sum over
# boards x squares of products that largely cancel, so the 2^-9 rounding of every activation / gradient element the device
# stores in bf16 shows up amplified, as noise around the true gradient.  precision="bf16" is the SAME fp32 graph with those
# tensors rounded at the same places: it measures what bf16 storage alone costs.  (Its roundings cannot be reproduced one
# by one: a single differing rounding perturbs ~1000 sums of the next layer, so after two or three layers the rounding
# decisions of any two implementations are independent.)  The bar: per tensor the device's distance to the fp32 gradient is
# no larger than NOISE_X times the bf16 oracle's own distance to it (+ NOISE_ABS), it points the same way (FP32_COS), and the
# two bf16 results are as close to each other as two independent roundings are (PAIR_X).  Losses: LOSS_TOL vs fp32.
NOISE_X, NOISE_ABS, PAIR_X, FP32_COS, LOSS_TOL = 2.5, 0.02, 2.0, 0.95, 2e-2


@pytest.mark.parametrize("filters,blocks,n", [(64, 1, 16), (48, 2, 33), (128, 2, 24), (256, 2, 16), (256, 3, 200)])
def test_train_step_against_autograd_oracle(filters, blocks, n):
    from chessbot_b200.train import Trainer
    from oracle import ref_model, ref_train
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    w = ref_model.init_weights(filters, blocks, seed=filters + blocks, random_bn=True)
    images, z, pi = _batch(n, seed=filters)
    lr = 0.05
    tr = Trainer(w, filters, blocks, n)
    tr.forward_backward(images, z, pi)
    g = tr.get_gradients()
    tr.apply(lr)
    losses = tr.losses()
    new_w = tr.get_weights()
    ref_losses, ref_new_w, _, ref_g = ref_train.train_step(w, None, images, z, pi, lr, device="cuda", precision="bf16")
    fp32_losses, _, _, fp32_g = ref_train.train_step(w, None, images, z, pi, lr, device="cuda", precision="fp32")
    for k in ("value", "policy", "l2"):
        assert abs(losses[k] - ref_losses[k]) <= 5e-3 * max(1.0, abs(ref_losses[k])), (k, losses, ref_losses)
        assert abs(losses[k] - fp32_losses[k]) <= LOSS_TOL * max(1.0, abs(fp32_losses[k])), (k, losses, fp32_losses)
    rows = {}
    for k in w:
        a, b, c = g[k], ref_g[k], fp32_g[k]
        if k.endswith(("bn", "bn1", "bn2")):
            a, b, c = a[:2], b[:2], c[:2]
        rows[k] = (round(_rel(a, c), 4), round(_rel(b, c), 4), round(_rel(a, b), 4), round(_cos(a, c), 5))
    worst = max(rows.items(), key=lambda kv: kv[1][0])
    print(f"\ntrain step {filters}x{blocks} batch {n}: losses {losses}; worst gradient (device vs fp32, bf16 oracle vs fp32, device vs "
          f"bf16 oracle, cos device/fp32): {worst}")
    big = {k: v for k, v in rows.items() if w[k].size > 8}        # value.bn / policy.bn: 2 / 4 numbers, covered by the update check
    bad = {k: v for k, v in big.items() if v[0] > NOISE_X * v[1] + NOISE_ABS or v[3] < FP32_COS or v[2] > PAIR_X * max(v[0], v[1]) + NOISE_ABS}
    assert not bad, bad
    # the SGD(momentum, nesterov) + l2 update and the BN moving statistics (moving*0.99 + batch*0.01)
    # (checked against the update the oracle's optimizer makes from the DEVICE's gradients: the optimizer itself is exact fp32)
    own_w, _ = ref_train.sgd_step(w, g, None, {k: (new_w[k][2], new_w[k][3]) for k in w if k.endswith(("bn", "bn1", "bn2"))}, lr)
    for k in w:
        if k.endswith(("bn", "bn1", "bn2")):
            assert np.allclose(new_w[k][:2], own_w[k][:2], rtol=1e-5, atol=1e-6), k
            assert np.allclose(new_w[k][2:], ref_new_w[k][2:], rtol=2e-3, atol=2e-4), k      # moving statistics vs the oracle's batch statistics
        else:
            assert np.allclose(new_w[k], own_w[k], rtol=1e-5, atol=1e-6), k


def test_training_reduces_the_loss_and_weights_serve_the_search():
    """a few steps on one batch lower the loss; the trained weights load into the inference path (DeviceNetwork)."""
    from chessbot_b200.model import DeviceNetwork
    from chessbot_b200.train import Trainer
    from oracle import ref_model
    filters, blocks, n = 64, 2, 32
    w = ref_model.init_weights(filters, blocks, seed=3)
    images, z, pi = _batch(n, seed=9)
    tr = Trainer(w, filters, blocks, n)
    first = tr.train_on_batch(images, z, pi, 0.02)
    for _ in range(15):
        last = tr.train_on_batch(images, z, pi, 0.02)
    assert np.isfinite(last["loss"]) and last["loss"] < first["loss"] - 0.5 and last["policy"] < first["policy"], (first, last)
    trained = tr.get_weights()
    net = DeviceNetwork(trained, filters, blocks)
    v_ref, p_ref = ref_model.forward(trained, images[:4], device="cuda")
    out = net(images[:1].astype(np.float32))
    assert abs(float(np.asarray(out["value_head"]).ravel()[0]) - float(v_ref[0])) < 5e-2


def test_checkpoint_resume_and_error_paths(tmp_path):
    """Trainer.save / load restore weights, optimizer momentum and step (train.py:124-136 resumes from its Keras checkpoint);
    the C ABI rejects what it cannot train instead of falling back."""
    from chessbot_b200._lib import lib
    from chessbot_b200.train import Trainer
    from oracle import ref_model
    filters, blocks, n = 64, 1, 8
    w = ref_model.init_weights(filters, blocks, seed=11)
    images, z, pi = _batch(n, seed=11)
    a = Trainer(w, filters, blocks, n)
    a.train_on_batch(images, z, pi, 0.05)
    a.train_on_batch(images, z, pi, 0.05)
    path = str(tmp_path / "ckpt.npz")
    a.save(path)
    b = Trainer(ref_model.init_weights(filters, blocks, seed=12), filters, blocks, n)
    b.load(path)
    assert b.steps == 2
    wa, wb, va, vb = a.get_weights(), b.get_weights(), a.get_velocity(), b.get_velocity()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa) and all(np.array_equal(va[k], vb[k]) for k in va)   # restored exactly
    # the next step of both agrees to rounding (batch statistics are summed with atomics: order-dependent in the last bit, and
    # one flipped bf16 rounding moves the whole gradient by its rounding noise, see the note above NOISE_X)
    la, lb = a.train_on_batch(images, z, pi, 0.05), b.train_on_batch(images, z, pi, 0.05)
    assert abs(la["loss"] - lb["loss"]) < 5e-3
    wa, wb = a.get_weights(), b.get_weights()
    assert all(_rel(wa[k], wb[k]) < 2e-2 for k in wa)
    with pytest.raises(RuntimeError):
        a.forward_backward(images[:4], z[:4], pi[:4])                  # a trainer is built for one batch size
    with pytest.raises(RuntimeError):
        Trainer(ref_model.init_weights(50, 1, seed=1), 50, 1, 4)       # filters must be a multiple of 4
    with pytest.raises(ValueError):
        a.forward_backward(images[..., :100], z, pi)
    assert lib.cb_train_set_tensor(a._h, b"nonsense.w", 0, None, 0, None) < 0


def test_selfplay_to_training_cycle_runs_end_to_end():
    """actors -> replay buffer -> trainer -> new network for the actors (train.py's cycle) on the device path, toy scale:
    samples have the reference's format, losses are finite, and the trained weights differ from the initial ones."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("demo", os.path.join(os.path.dirname(__file__), "..", "scripts", "selfplay_train_demo.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    lines = []
    hist = demo.run(net="2x32", games=12, iters=2, batch=16, steps_per_iter=2, sims=(4, 8), max_game_length=20, log=lines.append)
    assert len(hist) == 2 and hist[-1][0] >= 16, lines
    logs = hist[-1][1]
    assert logs is not None and all(np.isfinite(v) for v in logs.values()) and 5.0 < logs["policy"] < 9.5, logs


def test_config5_full_size_step_against_fp32_oracle():
    """BASELINE config 5 at full size: 20x256 net, batch 4096, one step against the torch-fp32 autograd oracle (on the GPU: it
    needs seconds there).  41 layers of bf16 activations and gradients: the bar is the loss to LOSS_TOL and, per tensor, the
    direction of the gradient (cosine) with no scale bias; plus exact size-independent properties of the device's own
    gradients: the policy-logit gradient of every sample sums to zero, so every row of the policy Dense gradient sums to ~0."""
    from chessbot_b200.train import Trainer
    from oracle import ref_model, ref_train
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    filters, blocks, n = 256, 20, 4096
    w = ref_model.init_weights(filters, blocks, seed=20)
    images, z, pi = _batch(256, seed=5)
    reps = n // 256
    rng = np.random.default_rng(1)
    images = np.concatenate([images[rng.permutation(256)] for _ in range(reps)])     # 4096 samples drawn from 256 real positions
    z, pi = ref_train.synthetic_targets(n, 6)
    tr = Trainer(w, filters, blocks, n)
    tr.forward_backward(images, z, pi)
    g = tr.get_gradients()
    tr.apply(0.01)
    losses = tr.losses()
    ref_losses, ref_g, _, _, _ = ref_train.forward_backward(w, images, z, pi, device="cuda")
    for k in ("value", "policy", "l2"):
        assert abs(losses[k] - ref_losses[k]) <= LOSS_TOL * max(1.0, abs(ref_losses[k])), (k, losses, ref_losses)
    rows = {}
    for k in w:
        a, b = (g[k][:2], ref_g[k][:2]) if k.endswith(("bn", "bn1", "bn2")) else (g[k], ref_g[k])
        rows[k] = (round(_rel(a, b), 3), round(_cos(a, b), 4), round(_proj(a, b), 3))
    worst = min(rows.items(), key=lambda kv: kv[1][1])
    print(f"\nconfig 5 full size (20x256, batch 4096): losses {losses} (oracle {ref_losses}); worst gradient cosine {worst}; "
          f"stem {rows['stem.w']}, res10.w1 {rows['res10.w1']}, res19.w2 {rows['res19.w2']}, policy.fc {rows['policy.fc']}")
    big = {k: v for k, v in rows.items() if w[k].size > 8}
    # measured: relative L2 error 2-5 %, cosine >= 0.9989, projection 0.998-1.003 (the large batch averages the rounding noise)
    bad = {k: v for k, v in big.items() if v[0] > 0.10 or v[1] < 0.995 or abs(v[2] - 1) > 0.02}
    assert not bad, bad
    row_sums = np.abs(g["policy.fc"].sum(axis=1))
    assert float(row_sums.max()) <= 1e-4 * float(np.abs(g["policy.fc"]).sum(axis=1).max()) + 1e-6


def test_training_mode_and_inference_path_agree_once_the_running_statistics_settle():
    """Train for a while, then freeze the weights (lr 0, momentum 0) and keep stepping on the SAME batch: the BN running statistics
    converge to that batch's statistics (moving = moving * 0.99 + batch * 0.01), so the INFERENCE kernels (folded running
    statistics, one persistent tower launch) on the trained weights must reproduce the training-mode outputs of the batch:
    training-mode BN, the moving-average rule, get_weights and the inference path are consistent with each other."""
    from chessbot_b200.model import DeviceNetwork
    from chessbot_b200.train import Trainer
    from oracle import ref_model
    filters, blocks, n = 48, 2, 64
    images, _, pi = _batch(n, seed=21)
    z = np.sign(images[:, 0, 0, 112] - 0.5).astype(np.float32)
    tr = Trainer(ref_model.init_weights(filters, blocks, seed=2), filters, blocks, n)
    first = tr.train_on_batch(images, z, pi, 0.02)
    for _ in range(150):
        last = tr.train_on_batch(images, z, pi, 0.02)
    assert np.isfinite(last["loss"]) and last["loss"] < first["loss"], (first, last)
    before = tr.get_weights()
    for _ in range(1200):                                  # 0.99^1200 = 6e-6 of the old running statistics left
        tr.forward_backward(images, z, pi)
        tr.apply(0.0, momentum=0.0)
    after = tr.get_weights()
    assert all(np.array_equal(before[k], after[k]) for k in before if "bn" not in k)          # lr 0, momentum 0: weights frozen
    train_mode = tr.values()
    value, _ = DeviceNetwork(after, filters, blocks, max_batch=n).predict(images)
    assert float(np.abs(value.cpu().numpy() - train_mode).max()) < 3e-2, float(np.abs(value.cpu().numpy() - train_mode).max())
    assert abs(tr.losses()["value"] - float(np.mean((train_mode - z) ** 2))) < 1e-4


def test_device_network_fit_and_save_like_the_keras_model_in_train_py(tmp_path):
    """train.py:150-164 calls model.fit(x, {"value_head": yv, "policy_head": yp}, initial_epoch, epochs, callbacks) and model.save on
    its Keras model; the DeviceNetwork takes the same calls: two minibatches of 32 per epoch (Keras' default batch size on
    config.batch_size = 64 samples), learning rate from the scheduler callback, trained weights used for inference afterwards."""
    from chessbot_b200.model import DeviceNetwork, load_network
    from oracle import ref_model
    images, z, pi = _batch(64, seed=31)
    z = np.sign(images[:, 0, 0, 112] - 0.5).astype(np.float64)
    net = DeviceNetwork(ref_model.init_weights(48, 1, seed=4), 48, 1, max_batch=64)
    before = net.predict(images)[0].cpu().numpy()
    seen = []

    class LrCallback:                                   # tf.keras.callbacks.LearningRateScheduler's interface
        def schedule(self, epoch, lr):
            seen.append((epoch, lr))
            return 0.02

    hists = []
    for i in range(1, 31):
        hists.append(net.fit(images, {"value_head": z, "policy_head": pi.astype(np.float64)}, initial_epoch=(i - 1), epochs=i,
                             callbacks=[LrCallback()], rng=np.random.default_rng(i)))
    assert [e for e, _ in seen] == list(range(30)) and seen[0][1] == 0.1 and seen[1][1] == 0.02
    assert net._trainer.steps == 60 and len(hists[-1].history["loss"]) == 1
    # training-mode losses fall (the inference outputs follow once the BN running statistics, momentum 0.99, have caught up:
    # test_training_mode_and_inference_path_agree_once_the_running_statistics_settle)
    assert hists[-1].history["value_head_loss"][0] < 0.8 * hists[0].history["value_head_loss"][0]
    assert hists[-1].history["loss"][0] < hists[0].history["loss"][0]
    after = net.predict(images)[0].cpu().numpy()
    assert float(np.abs(after - before).max()) > 1e-3                         # the network evaluates with the trained weights
    path = str(tmp_path / "model.npz")
    net.save(path)
    again = load_network(path, max_batch=64)
    assert np.allclose(again.predict(images)[0].cpu().numpy(), after, atol=1e-6)


@pytest.mark.parametrize("filters,blocks,n", [(64, 0, 32), (128, 0, 24), (64, 1, 32)])
def test_shallow_network_gradients_against_the_bf16_rounding_oracle(filters, blocks, n):
    """in a deep tower the device and the bf16-rounding oracle drift apart like two independent bf16 implementations
    (a few per cent on a weight gradient), which could hide a defect of ONE kernel.  With no residual block (stem + heads) or a single
    one nothing accumulates: BN forward / backward, the stem and 256-wide weight-gradient paths, the data gradient, the skip path and the
    heads' backward are each compared tensor by tensor with the oracle that rounds where the device rounds.  Measured: 64x0 agrees to 5e-4 on
    every tensor; at 128x0 the policy branch agrees to 3e-4 and the value branch to 1.8e-2 (its gradient is proportional to value - z with
    |value| ~ 0.05 at random init, so the forward pass's 1e-3 absolute drift is a per-cent relative one); one block: 2-3e-2.  Bars: 2.5e-2 / 4e-2
    and cosine >= 0.999 - a kernel that dropped a term or mis-indexed a tensor sits far outside."""
    from chessbot_b200.train import Trainer
    from oracle import ref_model, ref_train
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    w = ref_model.init_weights(filters, blocks, seed=11 + filters + blocks, random_bn=True)
    images, z, pi = _batch(n, seed=3 + filters)
    tr = Trainer(w, filters, blocks, n)
    tr.forward_backward(images, z, pi)
    g = tr.get_gradients()
    losses = tr.losses()
    ref_losses, _, _, ref_g = ref_train.train_step(w, None, images, z, pi, 0.05, device="cuda", precision="bf16")
    rows = {}
    for k in w:
        a, b = g[k], ref_g[k]
        if k.endswith(("bn", "bn1", "bn2")):
            a, b = a[:2], b[:2]
        rows[k] = (round(_rel(a, b), 5), round(_cos(a, b), 6))
    print(f"\nshallow {filters}x{blocks} batch {n}: device vs bf16-rounding oracle (rel L2, cos): {rows}")
    for k in ("value", "policy"):
        assert abs(losses[k] - ref_losses[k]) <= 2e-3 * max(1.0, abs(ref_losses[k])), (k, losses, ref_losses)
    bar = 2.5e-2 if blocks == 0 else 4e-2
    bad = {k: v for k, v in rows.items() if w[k].size > 8 and (v[0] > bar or v[1] < 0.999)}
    assert not bad, bad
