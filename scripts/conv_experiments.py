"""Times one conv layer of the tower kernel (a one-layer plan through cb_debug_conv_layer) and the whole tower."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chessbot_b200._lib import lib
from chessbot_b200.engine import Engine, nhwc_to_act
from chessbot_b200.model import keras_init_weights

boards = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
filters = int(sys.argv[2]) if len(sys.argv) > 2 else 256
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 20
eng = Engine.get(0)
eng.load_network(keras_init_weights(filters, blocks, seed=0), filters, blocks, max_batch=boards)
cpad = (filters + 63) // 64 * 64
x = nhwc_to_act(torch.randn(boards, 8, 8, cpad, device="cuda"))
sk = nhwc_to_act(torch.randn(boards, 8, 8, cpad, device="cuda"))
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = torch.zeros_like(x)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, layer, skip in (("conv", 1, None), ("conv+skip", 2, sk)):
    us = timed(lambda: lib.cb_debug_conv_layer(eng._h, layer, C.c_void_p(x.data_ptr()), C.c_void_p(skip.data_ptr()) if skip is not None else None,
                                               C.c_void_p(out.data_ptr()), boards, 0, st), 50)
    fl = 2 * 64 * 9 * cpad * cpad * boards
    print(f"{name:10s} {us:8.1f} us  {fl / us / 1e6:8.1f} TFLOP/s (one-layer plan)")
planes = nhwc_to_act(torch.randn(boards, 8, 8, 128, device="cuda"))
us = timed(lambda: eng.forward(planes, boards, want_bf16=False, want_f32=False), 20)
fl = 2 * 64 * 9 * filters * (119 + 2 * blocks * filters) * boards
print(f"tower {blocks}x{filters} @ {boards} boards (+ heads): {us:8.1f} us  {fl / us / 1e6:8.1f} TFLOP/s algorithmic")
