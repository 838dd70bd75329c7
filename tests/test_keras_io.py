"""Checkpoint interchange: the minimal HDF5 reader / writer of chessbot_b200.keras_io and the Keras layer walk.
The fixture tests/golden/keras_2x16.keras is a `.keras` archive (zip: config.json, metadata.json, model.weights.h5 in the Keras 3
`layers/<name>/vars/<i>` layout) written by keras_io.save_keras_weights — a file saved by real Keras is not obtainable in the build
container (no TensorFlow / Keras / h5py), which DESIGN.md states."""
import os
import struct
import zipfile

import numpy as np
import pytest

from conftest import GOLDEN


def _weights(filters=16, blocks=2, seed=3):
    from chessbot_b200.model import keras_init_weights
    w = keras_init_weights(filters, blocks, seed=seed)
    rng = np.random.default_rng(seed)
    for k in w:
        if k.endswith("bn") or ".bn" in k:
            w[k] = (w[k] + rng.normal(0, 0.1, w[k].shape)).astype(np.float32)
            w[k][3] = np.abs(w[k][3]) + 0.5
    return w


def test_h5_round_trip_nested_groups_and_shapes():
    from chessbot_b200 import keras_io
    rng = np.random.default_rng(0)
    tree = {"layers": {f"conv2d_{i}": {"vars": {"0": rng.normal(size=(3, 3, 5, 7)).astype(np.float32)}} for i in range(90)},
            "scalars": {"a": np.float32(2.5).reshape(()), "empty": np.zeros((0, 4), np.float32), "v": np.arange(6, dtype=np.float32)}}
    got = keras_io.read_h5(keras_io.write_h5(tree))
    assert len(got) == 93
    for i in range(90):
        assert np.array_equal(got[f"layers/conv2d_{i}/vars/0"], tree["layers"][f"conv2d_{i}"]["vars"]["0"])
    assert got["scalars/a"].shape == () and float(got["scalars/a"]) == 2.5
    assert got["scalars/empty"].shape == (0, 4) and np.array_equal(got["scalars/v"], np.arange(6, dtype=np.float32))


def test_h5_file_structure_follows_the_specification():
    """spot checks of the bytes another HDF5 library keys on: signature, 8-byte offsets, root symbol-table entry, SNOD / TREE / HEAP
    signatures where the superblock and the object header say they are"""
    from chessbot_b200 import keras_io
    data = keras_io.write_h5({"g": {"x": np.ones((2, 3), np.float32)}})
    assert data[:8] == b"\x89HDF\r\n\x1a\n" and data[8] == 0 and data[13] == 8 and data[14] == 8
    eof = struct.unpack_from("<Q", data, 40)[0]
    assert eof == len(data)
    root_hdr, cache_type = struct.unpack_from("<QI", data, 64)
    btree, heap = struct.unpack_from("<QQ", data, 80)
    assert cache_type == 1 and data[btree:btree + 4] == b"TREE" and data[heap:heap + 4] == b"HEAP"
    assert data[root_hdr] == 1                                    # version 1 object header
    snod = struct.unpack_from("<Q", data, btree + 32)[0]
    assert data[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", data, snod + 6)[0] == 1
    assert all(a % 8 == 0 for a in (root_hdr, btree, heap, snod))


def test_reads_version2_object_headers_with_link_messages():
    """a hand-built new-style file (superblock 2, OHDR headers, compact links) as h5py writes with libver='latest'"""
    from chessbot_b200 import keras_io
    arr = np.arange(12, dtype="<f4").reshape(3, 4)
    buf = bytearray(48)

    def ohdr(msgs):
        body = b"".join(bytes([t]) + struct.pack("<HB", len(m), 0) + m for t, m in msgs)
        addr = len(buf)
        buf.extend(b"OHDR" + bytes([2, 0x00, len(body)]) + body + b"\0\0\0\0")
        return addr
    data_addr = len(buf)
    buf.extend(arr.tobytes())
    space = struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 3, 4)
    dtype = bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)
    ds = ohdr([(1, space), (3, dtype), (8, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes))])
    link = lambda name, a: bytes([1, 0x00, len(name)]) + name.encode() + struct.pack("<Q", a)
    linfo = struct.pack("<BBQQ", 0, 0, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)
    grp = ohdr([(2, linfo), (6, link("kernel", ds))])
    root = ohdr([(2, linfo), (6, link("dense", grp))])
    buf[:48] = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, 0xFFFFFFFFFFFFFFFF, len(buf), root) + b"\0\0\0\0"
    got = keras_io.read_h5(bytes(buf))
    assert list(got) == ["dense/kernel"] and np.array_equal(got["dense/kernel"], arr)


def test_keras_archive_round_trip_and_layer_walk(tmp_path):
    from chessbot_b200 import keras_io
    w = _weights()
    p = tmp_path / "model.keras"
    keras_io.save_keras_weights(str(p), w)
    with zipfile.ZipFile(p) as z:
        assert {"config.json", "metadata.json", "model.weights.h5"} <= set(z.namelist())
    got, filters, blocks = keras_io.load_keras(str(p))
    assert (filters, blocks) == (16, 2) and set(got) == set(w)
    for k in w:
        assert np.array_equal(got[k].reshape(w[k].shape), w[k]), k
    # plain .weights.h5 and the legacy "<layer>/<layer>/<variable>:0" layout
    q = tmp_path / "m.weights.h5"
    keras_io.save_keras_weights(str(q), w)
    assert all(np.array_equal(a.reshape(w[k].shape), w[k]) for k, a in keras_io.load_keras(str(q))[0].items())
    tree = keras_io.keras_layer_tree({**w, "value.conv": w["value.conv"].reshape(1, 1, 16, 1), "policy.conv": w["policy.conv"].reshape(1, 1, 16, 2),
                                      "value.fc2": w["value.fc2"].reshape(16, 1)}, 2)["layers"]
    legacy = {"model_weights": {n: {n: ({"kernel:0": v["vars"]["0"]} if len(v["vars"]) == 1 else
                                        {f"{nm}:0": v["vars"][str(i)] for i, nm in enumerate(("gamma", "beta", "moving_mean", "moving_variance"))})}
                                for n, v in tree.items()}}
    r = tmp_path / "legacy.h5"
    r.write_bytes(keras_io.write_h5(legacy))
    got2, f2, b2 = keras_io.load_keras(str(r))
    assert (f2, b2) == (16, 2) and all(np.array_equal(got2[k].reshape(w[k].shape), w[k]) for k in w)


def test_committed_fixture_loads():
    from chessbot_b200 import keras_io
    got, filters, blocks = keras_io.load_keras(os.path.join(GOLDEN, "keras_2x16.keras"))
    w = _weights()
    assert (filters, blocks) == (16, 2)
    for k in w:
        assert np.array_equal(got[k].reshape(w[k].shape), w[k]), k


def test_rejects_other_graphs():
    from chessbot_b200 import keras_io
    with pytest.raises(ValueError):
        keras_io.weights_from_keras_layers({"conv2d": [np.zeros((3, 3, 4, 4), np.float32)], "conv2d_1": [np.zeros((3, 3, 4, 4), np.float32)]})


@pytest.mark.gpu
def test_gpu_network_from_keras_archive_evaluates_like_the_oracle(tmp_path):
    """a `.keras` checkpoint -> DeviceNetwork (model.load_network) -> the device forward pass agrees with the torch-fp32 oracle"""
    import torch
    from chessbot_b200 import keras_io
    from chessbot_b200.model import load_network
    from oracle import ref_model
    rng = np.random.default_rng(5)
    images = rng.integers(0, 2, (8, 8, 8, 119)).astype(np.int16)
    w = ref_model.calibrate_bn(ref_model.init_weights(64, 2, seed=2), images, device="cuda")
    p = tmp_path / "model.keras"
    keras_io.save_keras_weights(str(p), w)
    net = load_network(str(p), max_batch=8)
    value, logits = net.predict(images.astype(np.float32))
    torch.cuda.synchronize()
    v_ref, p_ref = ref_model.forward(w, images, device="cuda")
    assert np.abs(value.cpu().numpy() - v_ref).max() <= 2e-2 and np.abs(logits.cpu().numpy() - p_ref).max() <= 2e-2
