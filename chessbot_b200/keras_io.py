"""Checkpoint interchange with the reference (train.py:88-101,124-136, init.py:44): reads the weights of a Keras
`model.save("….keras")` archive (or a `.weights.h5` / legacy `.h5` file) into the tensor names of include/chessbot_b200.h, and writes a
`.weights.h5` in the layout Keras 3 `model.load_weights()` expects — without TensorFlow, Keras or h5py, none of which this image has.

A `.keras` file is a zip archive holding `config.json`, `metadata.json` and `model.weights.h5`; the HDF5 file keeps one group per layer
with its variables as `vars/0`, `vars/1`, … (Keras 3 `saving_lib`), or `<layer>/<variable>:0` datasets in the legacy layout.  Below is a
minimal HDF5 reader / writer for exactly what such files contain: version 0-3 superblocks, version 1 and 2 object headers, groups as
symbol tables (B-tree v1 + local heap) or compact link messages, little-endian float / integer datasets stored contiguous, compact or
chunked without filters.  Written from the published HDF5 file-format specification; files produced by real Keras have not been
available in the build container (stated in DESIGN.md): the fixture under tests/golden/ is produced by `write_h5` below.

Mapping to the network of model.py:30-56 is by layer creation order and shapes (as scripts/keras_weights_bridge.py does on the
reference's side): the 3x3 convolutions and C-channel normalisations in order are stem, res0.w1 / bn1, res0.w2 / bn2, …; the 1x1
convolutions with 1 / 2 filters and the 1- / 2-channel normalisations are the heads; Dense "value_head" / "policy_head" the outputs, the
remaining Dense is value.fc1.  Keras layouts are kept: kernels HWIO, Dense [in, out], BN rows gamma, beta, moving_mean, moving_variance.
"""
import io
import json
import re
import struct
import zipfile

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# ----------------------------------------------------------------------------------------------- reader
class _H5Reader:
    def __init__(self, data: bytes):
        self.d = data
        base = data.find(_SIG)
        if base != 0:
            raise ValueError("not an HDF5 file")
        ver = data[8]
        if ver in (0, 1):
            self.so, self.sl = data[13], data[14]
            p = 24 if ver == 0 else 28
            p += 4 * self.so                                  # base, free-space, end-of-file, driver-info addresses
            self.root = self._u(p + self.so, self.so)         # root symbol-table entry: link name offset, then object header address
        elif ver in (2, 3):
            self.so, self.sl = data[9], data[10]
            self.root = self._u(12 + 3 * self.so, self.so)    # base, superblock extension, end of file, root object header
        else:
            raise ValueError(f"HDF5 superblock version {ver} not supported")
        if self.so != 8 or self.sl != 8:
            raise ValueError("only 8-byte offsets / lengths are supported")

    def _u(self, pos, n):
        return int.from_bytes(self.d[pos:pos + n], "little")

    # ---- object headers -> list of (type, bytes)
    def _messages(self, addr):
        d = self.d
        if d[addr:addr + 4] == b"OHDR":
            return self._messages_v2(addr)
        if d[addr] != 1:
            raise ValueError("unsupported object header")
        n_msgs, size = self._u(addr + 2, 2), self._u(addr + 8, 4)
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < n_msgs + 64:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end:
                mtype, msize = self._u(pos, 2), self._u(pos + 2, 2)
                body = d[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    blocks.append((self._u_b(body, 0, 8), self._u_b(body, 8, 8)))
                elif mtype != 0:
                    out.append((mtype, body))
        return out

    def _messages_v2(self, addr):
        d = self.d
        flags = d[addr + 5]
        pos = addr + 6
        if flags & 0x20:
            pos += 16
        if flags & 0x10:
            pos += 4
        nsz = 1 << (flags & 3)
        chunk0 = self._u(pos, nsz)
        pos += nsz
        track = bool(flags & 0x04)
        blocks, out = [(pos, chunk0)], []
        while blocks:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 4 + (2 if track else 0) <= end:
                mtype, msize = d[pos], self._u(pos + 1, 2)
                pos += 4 + (2 if track else 0)
                body = d[pos:pos + msize]
                pos += msize
                if mtype == 0x10:
                    a, ln = self._u_b(body, 0, 8), self._u_b(body, 8, 8)
                    blocks.append((a + 4, ln - 8))             # OCHK signature in front, checksum behind
                elif mtype != 0:
                    out.append((mtype, body))
        return out

    @staticmethod
    def _u_b(b, pos, n):
        return int.from_bytes(b[pos:pos + n], "little")

    # ---- groups
    def _children(self, msgs):
        """{name: object header address}"""
        out = {}
        for mtype, body in msgs:
            if mtype == 0x11:                                  # symbol table: B-tree v1 of group nodes + local heap with the names
                btree, heap = self._u_b(body, 0, 8), self._u_b(body, 8, 8)
                if self.d[heap:heap + 4] != b"HEAP":
                    raise ValueError("bad local heap")
                heap_data = self._u(heap + 24, 8)
                self._walk_group_btree(btree, heap_data, out)
            elif mtype == 0x06:                                # link message (compact new-style group)
                ver, fl = body[0], body[1]
                p = 2
                ltype = 0
                if fl & 0x08:
                    ltype = body[p]; p += 1
                if fl & 0x04:
                    p += 8
                if fl & 0x10:
                    p += 1
                lsz = 1 << (fl & 3)
                ln = self._u_b(body, p, lsz); p += lsz
                name = body[p:p + ln].decode(); p += ln
                if ltype == 0:
                    out[name] = self._u_b(body, p, 8)
            elif mtype == 0x02:                                # link info: dense storage is not handled
                ver, fl = body[0], body[1]
                p = 2 + (8 if fl & 1 else 0)
                fheap = self._u_b(body, p, 8)
                if fheap != _UNDEF:
                    raise ValueError("HDF5 group with dense link storage (fractal heap) is not supported")
        return out

    def _walk_group_btree(self, addr, heap_data, out):
        d = self.d
        if d[addr:addr + 4] == b"SNOD":
            n = self._u(addr + 6, 2)
            for i in range(n):
                e = addr + 8 + i * 40
                name_off, obj = self._u(e, 8), self._u(e + 8, 8)
                s = heap_data + name_off
                out[d[s:d.index(b"\x00", s)].decode()] = obj
            return
        if d[addr:addr + 4] != b"TREE":
            raise ValueError("bad group B-tree node")
        used = self._u(addr + 6, 2)
        p = addr + 24 + 8                                      # signature, type, level, entries, two siblings, then key 0
        for _ in range(used):
            self._walk_group_btree(self._u(p, 8), heap_data, out)
            p += 16                                            # child address + next key

    # ---- datasets
    def _dataset(self, msgs):
        shape = dtype = layout = None
        for mtype, body in msgs:
            if mtype == 0x01:
                ver, rank = body[0], body[1]
                p = 8 if ver == 1 else 4
                shape = tuple(self._u_b(body, p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                cls, size = body[0] & 0x0F, self._u_b(body, 4, 4)
                if body[1] & 1:
                    raise ValueError("big-endian HDF5 data is not supported")
                if cls == 1:
                    dtype = np.dtype(f"<f{size}")
                elif cls == 0:
                    dtype = np.dtype(f"<{'i' if body[1] & 0x08 else 'u'}{size}")
                else:
                    dtype = np.dtype(f"V{size}")
            elif mtype == 0x08:
                layout = body
        if shape is None or dtype is None or layout is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        nbytes = count * dtype.itemsize
        ver, cls = layout[0], layout[1]
        if ver != 3 and ver != 4:
            raise ValueError(f"data layout message version {ver} not supported")
        if cls == 0:
            raw = layout[4:4 + self._u_b(layout, 2, 2)]
        elif cls == 1:
            a = self._u_b(layout, 2, 8)
            raw = b"\x00" * nbytes if a == _UNDEF else self.d[a:a + nbytes]
        elif cls == 2 and ver == 3:
            raw = self._read_chunked(layout, shape, dtype)
        else:
            raise ValueError("unsupported dataset layout")
        if dtype.kind == "V":
            return None
        return np.frombuffer(raw[:nbytes], dtype=dtype).reshape(shape).copy()

    def _read_chunked(self, layout, shape, dtype):
        rank1 = layout[2]
        btree = self._u_b(layout, 3, 8)
        cdims = [self._u_b(layout, 11 + 4 * i, 4) for i in range(rank1)][:-1]
        out = np.zeros(shape, dtype=dtype)
        if btree == _UNDEF:
            return out.tobytes()

        def walk(addr):
            d = self.d
            if d[addr:addr + 4] != b"TREE" or d[addr + 4] != 1:
                raise ValueError("bad chunk B-tree node")
            level, used = d[addr + 5], self._u(addr + 6, 2)
            p = addr + 24
            ksz = 8 + 8 * rank1
            for _ in range(used):
                csize, mask = self._u(p, 4), self._u(p + 4, 4)
                offs = [self._u(p + 8 + 8 * i, 8) for i in range(rank1 - 1)]
                child = self._u(p + ksz, 8)
                if level:
                    walk(child)
                else:
                    if mask or csize != int(np.prod(cdims)) * dtype.itemsize:
                        raise ValueError("filtered (compressed) HDF5 chunks are not supported")
                    chunk = np.frombuffer(d[child:child + csize], dtype=dtype).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
                p += ksz + 8
        walk(btree)
        return out.tobytes()

    def datasets(self):
        """{path: ndarray} of every numeric dataset in the file"""
        out, seen = {}, set()

        def visit(addr, path):
            if addr in seen:
                return
            seen.add(addr)
            msgs = self._messages(addr)
            kids = self._children(msgs)
            if kids or any(t in (0x11, 0x02) for t, _ in msgs):
                for name, a in kids.items():
                    visit(a, f"{path}/{name}" if path else name)
                return
            arr = self._dataset(msgs)
            if arr is not None:
                out[path] = arr
        visit(self.root, "")
        return out


def read_h5(source):
    """{dataset path: ndarray} for an HDF5 file given as a path, bytes or a binary file object."""
    if isinstance(source, (bytes, bytearray)):
        data = bytes(source)
    elif hasattr(source, "read"):
        data = source.read()
    else:
        with open(source, "rb") as f:
            data = f.read()
    return _H5Reader(data).datasets()


# ----------------------------------------------------------------------------------------------- writer
class _H5Writer:
    """Version 0 superblock, version 1 object headers, one symbol-table node per group (up to 2 * leaf K entries), contiguous
    little-endian float32 datasets: the subset every HDF5 library reads."""
    LEAF_K = 64

    def __init__(self):
        self.buf = bytearray(b"\x00" * 96)                     # superblock (56 + 40 bytes), filled in at the end

    def _align(self):
        while len(self.buf) % 8:
            self.buf.append(0)

    def _alloc(self, blob):
        self._align()
        addr = len(self.buf)
        self.buf += blob
        return addr

    @staticmethod
    def _msg(mtype, body):
        body = body + b"\x00" * (-len(body) % 8)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def _header(self, msgs):
        data = b"".join(msgs)
        return self._alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(data)) + data)

    def dataset(self, arr):
        arr = np.asarray(arr, dtype="<f4")                     # (ascontiguousarray would turn a scalar into shape (1,))
        addr = self._alloc(arr.tobytes()) if arr.size else _UNDEF
        space = struct.pack("<BBBx4x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
        dtype = bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)                # fill value v2: allocate late, write at allocation, undefined
        layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
        return self._header([self._msg(1, space), self._msg(3, dtype), self._msg(5, fill), self._msg(8, layout)])

    def group(self, children):
        """children: {name: object header address}; returns (header address, btree address, heap address)"""
        names = sorted(children)                                # a symbol-table node keeps its entries sorted by name
        if len(names) > 2 * self.LEAF_K:
            raise ValueError("group too large for the minimal writer")
        heap_data = bytearray(b"\x00" * 8)                     # offset 0: the empty name of the first B-tree key
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode() + b"\x00"
            heap_data += b"\x00" * (-len(heap_data) % 8)
        free = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)                  # one free block to the end of the segment (offset 1 = last)
        data_addr = self._alloc(bytes(heap_data))
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free, data_addr))
        entries = b"".join(struct.pack("<QQII16x", offs[n], children[n], 0, 0) for n in names)
        entries += b"\x00" * (40 * (2 * self.LEAF_K - len(names)))
        snod = self._alloc(b"SNOD" + struct.pack("<BxH", 1, len(names)) + entries)
        last = offs[names[-1]] if names else 0
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, _UNDEF, _UNDEF) + struct.pack("<QQQ", 0, snod, last)
        node += b"\x00" * (16 * 2 * 16 - 16)                    # room for 2 * internal K entries, as libhdf5 lays nodes out
        btree = self._alloc(node)
        hdr = self._header([self._msg(0x11, struct.pack("<QQ", btree, heap))])
        return hdr, btree, heap

    def finish(self, root):
        hdr, btree, heap = root
        self._align()
        sb = _SIG + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, self.LEAF_K, 16, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        self.buf[:len(sb)] = sb
        return bytes(self.buf)


def write_h5(tree):
    """bytes of an HDF5 file for a nested dict {name: dict | ndarray} (float32 datasets)."""
    w = _H5Writer()

    def build(node):
        kids = {k: (w.dataset(v) if not isinstance(v, dict) else build(v)[0]) for k, v in node.items()}
        return w.group(kids)
    return w.finish(build(tree))


# ----------------------------------------------------------------------------------------------- Keras <-> tensor names
_BN_VARS = ("gamma", "beta", "moving_mean", "moving_variance")


def _layer_key(name):
    m = re.fullmatch(r"(.*?)(?:_(\d+))?", name)
    return (m.group(1), int(m.group(2)) if m.group(2) else 0)


def _collect_layers(datasets):
    """{layer name: [variables in Keras order]} from either HDF5 layout"""
    layers = {}
    for path, arr in datasets.items():
        parts = path.split("/")
        if "vars" in parts:                                     # Keras 3: <…>/<layer>/vars/<index>
            i = parts.index("vars")
            if i == 0 or "optimizer" in parts[:i]:
                continue
            layers.setdefault(parts[i - 1], {})[int(parts[-1])] = arr
        elif ":" in parts[-1] and "optimizer" not in parts[0]:  # legacy: model_weights/<layer>/<layer>/<variable>:0
            var = parts[-1].split(":")[0]
            order = {"kernel": 0, "gamma": 0, "beta": 1, "moving_mean": 2, "moving_variance": 3}
            if var in order:
                layers.setdefault(parts[-2], {})[order[var]] = arr
    return {k: [v[i] for i in sorted(v)] for k, v in layers.items()}


def weights_from_keras_layers(layers):
    """(weights dict in chessbot_b200 tensor names, filters, blocks) from {layer name: [variables]} of model.generate_model()."""
    convs3, bns, out = [], [], {}
    for name in sorted(layers, key=_layer_key):
        v = layers[name]
        if len(v) == 1 and v[0].ndim == 4:
            k = v[0]
            if k.shape[0] == 3:
                convs3.append(k)
            elif k.shape[-1] == 1:
                out["value.conv"] = k
            elif k.shape[-1] == 2:
                out["policy.conv"] = k
        elif len(v) == 4 and all(a.ndim == 1 for a in v):
            c = v[0].shape[0]
            row = np.stack(v).astype(np.float32)
            if c == 1:
                out["value.bn"] = row
            elif c == 2:
                out["policy.bn"] = row
            else:
                bns.append(row)
        elif len(v) == 1 and v[0].ndim == 2:
            if name == "value_head":
                out["value.fc2"] = v[0]
            elif name == "policy_head":
                out["policy.fc"] = v[0]
            else:
                out["value.fc1"] = v[0]
    if len(convs3) != len(bns) or len(convs3) % 2 != 1:
        raise ValueError(f"not the graph of model.generate_model(): {len(convs3)} 3x3 convolutions, {len(bns)} wide normalisations")
    out["stem.w"], out["stem.bn"] = convs3[0], bns[0]
    blocks = (len(convs3) - 1) // 2
    for i in range(blocks):
        out[f"res{i}.w1"], out[f"res{i}.bn1"] = convs3[1 + 2 * i], bns[1 + 2 * i]
        out[f"res{i}.w2"], out[f"res{i}.bn2"] = convs3[2 + 2 * i], bns[2 + 2 * i]
    missing = [k for k in ("value.conv", "value.bn", "value.fc1", "value.fc2", "policy.conv", "policy.bn", "policy.fc") if k not in out]
    if missing:
        raise ValueError(f"head tensors not found: {missing}")
    return {k: np.asarray(a, dtype=np.float32) for k, a in out.items()}, int(convs3[0].shape[-1]), blocks


def load_keras(path):
    """(weights, filters, blocks) from `model.save("x.keras")`, `model.save_weights("x.weights.h5")` or a legacy `.h5` file of the
    reference's model (train.py:124-136 loads checkpoints/keras/model.keras)."""
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            name = next((n for n in z.namelist() if n.endswith("model.weights.h5")), None)
            if name is None:
                raise ValueError(f"{path}: no model.weights.h5 in the archive")
            data = z.read(name)
    else:
        with open(path, "rb") as f:
            data = f.read()
    return weights_from_keras_layers(_collect_layers(read_h5(data)))


def keras_layer_tree(weights, blocks):
    """Keras 3 `layers/<auto name>/vars/<i>` tree of model.generate_model() (layer auto-names in creation order, model.py:30-56)."""
    layers, n_conv, n_bn = {}, 0, 0

    def name(base, n):
        return base if n == 0 else f"{base}_{n}"

    def conv(key):
        nonlocal n_conv
        layers[name("conv2d", n_conv)] = {"vars": {"0": weights[key]}}
        n_conv += 1

    def bn(key):
        nonlocal n_bn
        layers[name("batch_normalization", n_bn)] = {"vars": {str(i): weights[key][i] for i in range(4)}}
        n_bn += 1
    conv("stem.w"); bn("stem.bn")
    for i in range(blocks):
        conv(f"res{i}.w1"); bn(f"res{i}.bn1"); conv(f"res{i}.w2"); bn(f"res{i}.bn2")
    conv("value.conv"); bn("value.bn")
    layers["dense"] = {"vars": {"0": weights["value.fc1"]}}
    layers["value_head"] = {"vars": {"0": weights["value.fc2"]}}
    conv("policy.conv"); bn("policy.bn")
    layers["policy_head"] = {"vars": {"0": weights["policy.fc"]}}
    return {"layers": layers}


def save_keras_weights(path, weights, blocks=None):
    """Writes `path` (`….weights.h5`, or a `.keras` archive if the name ends in .keras) holding these weights in the Keras 3 layout
    for the reference's `model.load_weights()` / a `.keras` whose model.weights.h5 replaces the one of an archive saved by Keras."""
    blocks = sum(1 for k in weights if k.endswith(".w1")) if blocks is None else blocks
    w = {k: np.asarray(v, dtype=np.float32) for k, v in weights.items()}
    for k in ("value.conv", "policy.conv"):
        w[k] = w[k].reshape(1, 1, w[k].shape[-2] if w[k].ndim > 1 else w[k].size, -1)
    w["value.fc2"] = w["value.fc2"].reshape(-1, 1)
    data = write_h5(keras_layer_tree(w, blocks))
    if str(path).endswith(".keras"):
        meta = {"keras_version": "3", "date_saved": "", "note": "weights written by chessbot_b200.keras_io; config.json must come from Keras"}
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("metadata.json", json.dumps(meta))
            z.writestr("config.json", json.dumps({"class_name": "Functional", "note": "architecture: model.generate_model() of the reference"}))
            z.writestr("model.weights.h5", data)
    else:
        with open(path, "wb") as f:
            f.write(data)
