"""several GPUs in ONE process.  Kernel attributes (dynamic shared memory opt-in) are per device and every C-ABI entry
point must run on its context's device whatever the caller's current device is.  Needs two visible GPUs (skipped otherwise)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_devices_in_one_process_forward_and_search():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from chessbot_b200.board import Board
    from chessbot_b200.engine import EVAL_NET, Engine
    from chessbot_b200.model import DeviceNetwork, keras_init_weights
    w = keras_init_weights(64, 2, seed=4)
    boards = []
    for line in ("", "e2e4 e7e5", "d2d4 d7d5 c2c4", "g1f3 g8f6 c2c4 e7e6 b1c3"):
        b = Board()
        for u in line.split():
            b.push_uci(u)
        boards.append(b)
    recs = [b.export_record() for b in boards]
    outs = []
    for dev in (0, 1):
        torch.cuda.set_device(dev)
        net = DeviceNetwork(w, 64, 2, max_batch=64, device=dev)        # first tower launch on this device sets its kernel attributes
        eng = net.engine
        assert eng.device_index == dev
        pos, fi = eng.upload_records(recs)
        act, _ = eng.encode(pos, fi)
        v, lb, _ = eng.forward(act, len(recs))
        torch.cuda.synchronize(dev)
        assert v.device.index == dev
        outs.append((v.cpu(), lb.float().cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # entry points called while ANOTHER device is current still run on their own context's device
    torch.cuda.set_device(0)
    eng1 = Engine.get(1)
    torch.cuda.set_device(1)
    eng1.search_create(8, max_nodes=8 * 40 * 48, max_positions=8 * 300, max_sims=32)
    out = eng1.search(recs, 16, eval_mode=EVAL_NET)
    assert out["unfinished"] == 0 and out["root_N"].tolist() == [17] * len(recs)
    eng0 = Engine.get(0)                     # driven while device 1 is current: streams and buffers are taken from the engine's own device
    eng0.search_create(8, max_nodes=8 * 40 * 48, max_positions=8 * 300, max_sims=32)
    out0 = eng0.search(recs, 16, eval_mode=EVAL_NET)
    assert torch.cuda.current_device() == 1
    assert np.array_equal(out0["N"], out["N"])
