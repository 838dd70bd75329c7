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
