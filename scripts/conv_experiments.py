"""Times one 256-channel conv layer (1024 boards) under the CB_CONV_DEBUG variants; run one variant per process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chessbot_b200.engine import Engine, nhwc_to_act
from chessbot_b200.model import keras_init_weights
boards, filters = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 256
eng = Engine.get(0)
eng.load_network(keras_init_weights(filters, 1, seed=0), filters, 1, max_batch=boards)
x = nhwc_to_act(torch.randn(boards, 8, 8, filters, device="cuda"))
sk = nhwc_to_act(torch.randn(boards, 8, 8, filters, device="cuda"))
for name, layer, skip in (("conv", 1, None), ("conv+skip", 2, sk)):
    for _ in range(5): eng.debug_conv_layer(layer, x, skip, boards, filters)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    from chessbot_b200._lib import lib
    import ctypes as C
    out = torch.zeros_like(x)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    e0.record()
    for _ in range(reps):
        lib.cb_debug_conv_layer(eng._h, layer, C.c_void_p(x.data_ptr()), C.c_void_p(skip.data_ptr()) if skip is not None else None,
                                C.c_void_p(out.data_ptr()), boards, 0, st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 2 * 64 * 9 * filters * filters * boards
    print(f"CB_CONV_DEBUG={os.environ.get('CB_CONV_DEBUG','0'):>3s} {name:10s} {us:8.1f} us  {fl / us / 1e6:8.1f} TFLOP/s")
