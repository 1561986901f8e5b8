"""Keras-2.7 ``.h5`` weight files without h5py: reader and writer for the checkpoints of the reference
(``model.save_weights`` -> ``saved_weights/NeRF_model_epoch_NNN.h5``, src/UtilsFiles.py:153-164, src/NeRF.py:343-351).

File layout the reference's Keras writes for ``NeRF(keras.Model)`` holding two functional models:
    /                      attrs layer_names = [model, model_1], backend, keras_version
    /model                 attr  weight_names = [dense/kernel:0, dense/bias:0, ..., dense_10/bias:0]
    /model/dense/kernel:0  float32 (in, out)       ... dense, dense_1 ... dense_10 = the coarse network, creation order
    /model_1/dense_11/...                          ... dense_11 ... dense_21 = the fine network
Only the HDF5 subset those files use is handled: superblock version 0, version-1 object headers, old-style groups
(symbol-table message -> version-1 B-tree -> SNOD leaves + local heap), contiguous little-endian float datasets.
The writer emits the same structures with fixed-length string attributes (what ``h5py`` produces for ``np.bytes_``
arrays and Keras' ``load_attributes_from_hdf5_group`` decodes).  It is checked against this reader and, structurally,
against the file the reference trained (tests/test_h5weights.py); libhdf5 itself is not available offline.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INTERNAL_K = 4, 16                       # group B-tree ranks written into the superblock
_SNOD_SIZE = 8 + 2 * _LEAF_K * 40
_TREE_SIZE = 24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8


# ---------------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------------
class H5Reader:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        if self.buf[:8] != _SIG:
            raise ValueError(f"{path}: not an HDF5 file")
        if self.buf[8] != 0:
            raise ValueError(f"{path}: superblock version {self.buf[8]} not supported (expected 0)")
        self.so, self.sl = self.buf[13], self.buf[14]
        self.base = self._off(24)
        self.root = self._symbol_entry(24 + 4 * self.so)

    def _off(self, pos):
        return int.from_bytes(self.buf[pos:pos + self.so], "little")

    def _len(self, pos):
        return int.from_bytes(self.buf[pos:pos + self.sl], "little")

    def _symbol_entry(self, pos):
        return {"name_off": self._off(pos), "header": self._off(pos + self.so)}

    def _messages(self, addr):
        if self.buf[addr] != 1:
            raise ValueError("only version-1 object headers are supported")
        n_msgs = struct.unpack_from("<H", self.buf, addr + 2)[0]
        blocks = [(addr + 16, struct.unpack_from("<I", self.buf, addr + 8)[0])]
        out = []
        while blocks and len(out) < n_msgs:
            pos, remaining = blocks.pop(0)
            end = pos + remaining
            while pos + 8 <= end and len(out) < n_msgs:
                mtype, msize = struct.unpack_from("<HH", self.buf, pos)
                out.append((mtype, pos + 8, msize))
                if mtype == 0x0010:                                   # continuation block
                    blocks.append((self._off(pos + 8), self._len(pos + 8 + self.so)))
                pos += 8 + msize
        return out

    def _heap_name(self, heap_addr, off):
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        start = self._off(heap_addr + 8 + 2 * self.sl) + off
        return self.buf[start:self.buf.index(b"\x00", start)].decode()

    def _btree_entries(self, btree_addr, heap_addr):
        if self.buf[btree_addr:btree_addr + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        level = self.buf[btree_addr + 5]
        n = struct.unpack_from("<H", self.buf, btree_addr + 6)[0]
        pos = btree_addr + 8 + 2 * self.so
        entries = []
        for _ in range(n):
            pos += self.sl
            child = self._off(pos)
            pos += self.so
            if level > 0:
                entries += self._btree_entries(child, heap_addr)
                continue
            if self.buf[child:child + 4] != b"SNOD":
                raise ValueError("bad symbol-table node")
            epos = child + 8
            for _ in range(struct.unpack_from("<H", self.buf, child + 6)[0]):
                e = self._symbol_entry(epos)
                e["name"] = self._heap_name(heap_addr, e["name_off"])
                entries.append(e)
                epos += 2 * self.so + 24
        return entries

    def children(self, header_addr):
        for mtype, data, _ in self._messages(header_addr):
            if mtype == 0x0011:
                return self._btree_entries(self._off(data), self._off(data + self.so))
        return None

    def dataset(self, header_addr):
        shape = dtype = address = None
        for mtype, data, _ in self._messages(header_addr):
            if mtype == 0x0001:
                version, rank = self.buf[data], self.buf[data + 1]
                dims = data + (8 if version == 1 else 4)
                shape = tuple(self._len(dims + i * self.sl) for i in range(rank))
            elif mtype == 0x0003:
                if self.buf[data] & 0x0F != 1:
                    return None
                dtype = {4: "<f4", 8: "<f8", 2: "<f2"}[struct.unpack_from("<I", self.buf, data + 4)[0]]
            elif mtype == 0x0008:
                if self.buf[data] != 3 or self.buf[data + 1] != 1:
                    raise ValueError("only contiguous (layout v3) datasets are supported")
                address = self._off(data + 2)
        if shape is None or dtype is None or address is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        return np.frombuffer(self.buf, dtype=dtype, count=count, offset=self.base + address).reshape(shape).copy()

    def datasets(self):
        """{'/group/.../name': ndarray} of every float dataset."""
        out = {}

        def walk(header_addr, prefix):
            kids = self.children(header_addr)
            if kids is None:
                arr = self.dataset(header_addr)
                if arr is not None:
                    out[prefix] = arr
                return
            for e in kids:
                walk(e["header"], prefix + "/" + e["name"])
        walk(self.root["header"], "")
        return out


def _layer_index(name):
    return 0 if "_" not in name else int(name.rsplit("_", 1)[1])


def read_keras_weights(path):
    """[(model_group, [(layer_name, kernel, bias), ...]), ...] with layers in Keras creation order (dense, dense_1, ...)."""
    groups = {}
    for key, arr in H5Reader(path).datasets().items():
        parts = key.strip("/").split("/")
        if len(parts) < 3:
            continue
        model, layer, var = parts[0], parts[-2], parts[-1].split(":")[0]
        groups.setdefault(model, {}).setdefault(layer, {})[var] = arr.astype(np.float32)
    out = []
    for model in sorted(groups, key=_layer_index):
        layers = groups[model]
        out.append((model, [(n, layers[n]["kernel"], layers[n]["bias"]) for n in sorted(layers, key=_layer_index)]))
    return out


def load_flat_params(path):
    """One flat fp32 vector per network, [W0 (in,out) row-major, b0, W1, b1, ...]: (coarse, fine or None)."""
    models = read_keras_weights(path)
    flat = [np.concatenate([np.concatenate([k.reshape(-1), b.reshape(-1)]) for _, k, b in layers])
            for _, layers in models]
    if not flat:
        raise ValueError(f"{path}: no Dense weights found")
    return flat[0], (flat[1] if len(flat) > 1 else None)


# ---------------------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------------------
def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


def _string_attr(name, values, scalar=False):
    """Attribute message (version 1) holding fixed-length, null-padded ASCII strings."""
    values = [v.encode() if isinstance(v, str) else v for v in values]
    width = max([len(v) for v in values] + [1])
    dtype = struct.pack("<B3BI", 0x13, 0x01, 0, 0, width)                       # class 3 (string), null-pad, ASCII
    space = struct.pack("<BBB5x", 1, 0, 0) if scalar else struct.pack("<BBB5xQ", 1, 1, 0, len(values))
    nm = name.encode() + b"\x00"
    head = struct.pack("<BxHHH", 1, len(nm), len(dtype), len(space))
    data = b"".join(v.ljust(width, b"\x00") for v in values)
    return _message(0x000C, head + _pad8(nm) + _pad8(dtype) + _pad8(space) + data)


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)                                                 # superblock filled in at the end

    def alloc(self, data):
        addr = len(self.buf)
        self.buf += _pad8(bytes(data))
        return addr

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self.alloc(arr.tobytes())
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        dtype = bytes.fromhex("11201f000400000000002000170800177f000000")      # IEEE float32, little-endian
        fill = bytes([2, 2, 2, 1]) + struct.pack("<I", 0)                      # v2: late alloc, fill if set, default value
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        return self.alloc(_object_header([_message(0x0001, space), _message(0x0003, dtype, 1), _message(0x0005, fill, 1),
                                          _message(0x0008, layout)]))

    def group(self, children, attrs=()):
        """children: {name: object-header address}.  Returns (header address, btree address, heap address)."""
        names = sorted(children, key=lambda s: s.encode())                      # B-tree order = strcmp order
        heap = bytearray(8)                                                     # offset 0: the empty string
        offsets = {}
        for n in names:
            offsets[n] = len(heap)
            heap += _pad8(n.encode() + b"\x00")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 32) + bytes(16)                           # one free block (next = H5HL_FREE_NULL)
        heap_data = self.alloc(heap)
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        leaves = [names[i:i + 2 * _LEAF_K] for i in range(0, len(names), 2 * _LEAF_K)]
        if len(leaves) > 2 * _INTERNAL_K:
            raise ValueError("too many children for a single B-tree node")
        keys, kids = [0], []
        for leaf in leaves:
            node = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(leaf)))
            for n in leaf:
                node += struct.pack("<QQI4x16x", offsets[n], children[n], 0)
            node += bytes(_SNOD_SIZE - len(node))
            kids.append(self.alloc(node))
            keys.append(offsets[leaf[-1]])
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(kids), _UNDEF, _UNDEF))
        for i, kid in enumerate(kids):
            tree += struct.pack("<QQ", keys[i], kid)
        tree += struct.pack("<Q", keys[len(kids)])
        tree += bytes(_TREE_SIZE - len(tree))
        tree_addr = self.alloc(tree)
        header = self.alloc(_object_header([_message(0x0011, struct.pack("<QQ", tree_addr, heap_addr))] + list(attrs)))
        return header, tree_addr, heap_addr

    def finish(self, root):
        header, tree_addr, heap_addr = root
        sb = _SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQI4xQQ", 0, header, 1, tree_addr, heap_addr)     # root symbol-table entry (cached group)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_keras_weights(path, models, keras_version="2.7.0", backend="tensorflow"):
    """models: [(model_group_name, [(layer_name, kernel (in,out), bias), ...]), ...] -> Keras-2.7-style .h5 file."""
    w = _Writer()
    top = {}
    for model_name, layers in models:
        layer_groups, weight_names = {}, []
        for layer_name, kernel, bias in layers:
            kids = {"kernel:0": w.dataset(kernel), "bias:0": w.dataset(bias)}
            layer_groups[layer_name] = w.group(kids)[0]
            weight_names += [f"{layer_name}/kernel:0", f"{layer_name}/bias:0"]
        top[model_name] = w.group(layer_groups, [_string_attr("weight_names", weight_names)])[0]
    attrs = [_string_attr("layer_names", [m for m, _ in models]), _string_attr("backend", [backend], scalar=True),
             _string_attr("keras_version", [keras_version], scalar=True)]
    data = w.finish(w.group(top, attrs))
    with open(path, "wb") as f:
        f.write(data)


def save_flat_params(path, params_coarse, params_fine, shapes):
    """Write the two flat parameter vectors as the reference's checkpoint: layers ``dense .. dense_{L-1}`` in group
    ``model`` (coarse), ``dense_L .. dense_{2L-1}`` in ``model_1`` (fine).  shapes: [(in, out), ...] per Dense layer."""
    models, idx = [], 0
    for name, flat in (("model", params_coarse), ("model_1", params_fine)):
        if flat is None:
            continue
        flat = np.asarray(flat, dtype=np.float32).reshape(-1)
        layers, pos = [], 0
        for fan_in, fan_out in shapes:
            k = flat[pos:pos + fan_in * fan_out].reshape(fan_in, fan_out)
            pos += fan_in * fan_out
            b = flat[pos:pos + fan_out]
            pos += fan_out
            layers.append(("dense" if idx == 0 else f"dense_{idx}", k, b))
            idx += 1
        if pos != flat.size:
            raise ValueError("parameter vector does not match the layer shapes")
        models.append((name, layers))
    write_keras_weights(path, models)
