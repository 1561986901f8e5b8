"""Minimal pure-Python HDF5 reader for the Keras-2.7 weight files the reference writes
(model.save_weights -> `NeRF_model_epoch_NNN.h5`, src/UtilsFiles.py:153-164).  TEST INFRASTRUCTURE.

h5py is not installed in this image, so this module reads exactly the subset of HDF5 those files use:
superblock version 0, version-1 object headers (with continuation blocks), old-style groups (symbol-table
message -> v1 B-tree + local heap -> SNOD nodes), and contiguous little-endian float datasets.
It exists to pin the oracle against an output of the reference (trained weights + the PSNR the reference recorded
for them); the product never imports it.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        if self.buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        version = self.buf[8]
        if version != 0:
            raise ValueError(f"superblock version {version} not supported")
        self.so = self.buf[13]          # size of offsets
        self.sl = self.buf[14]          # size of lengths
        pos = 24
        self.base = self._off(pos)
        pos += 4 * self.so              # base, free-space, eof, driver-info addresses
        # root group symbol table entry
        self.root = self._symbol_entry(pos)

    # ---- primitive readers ----------------------------------------------------------------------------------------
    def _off(self, pos):
        return int.from_bytes(self.buf[pos:pos + self.so], "little")

    def _len(self, pos):
        return int.from_bytes(self.buf[pos:pos + self.sl], "little")

    def _symbol_entry(self, pos):
        name_off = self._off(pos)
        header = self._off(pos + self.so)
        cache_type = struct.unpack_from("<I", self.buf, pos + 2 * self.so)[0]
        scratch = pos + 2 * self.so + 8
        entry = {"name_off": name_off, "header": header, "cache_type": cache_type}
        if cache_type == 1:
            entry["btree"] = self._off(scratch)
            entry["heap"] = self._off(scratch + self.so)
        return entry

    # ---- object headers ----------------------------------------------------------------------------------------------
    def _messages(self, addr):
        """Yield (type, data_pos, size) of every message of a version-1 object header, following continuations."""
        if self.buf[addr] != 1:
            raise ValueError("only version-1 object headers are supported")
        n_msgs = struct.unpack_from("<H", self.buf, addr + 2)[0]
        size = struct.unpack_from("<I", self.buf, addr + 8)[0]
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            pos, remaining = blocks.pop(0)
            end = pos + remaining
            while pos + 8 <= end and len(out) < n_msgs:
                mtype, msize = struct.unpack_from("<HH", self.buf, pos)
                data = pos + 8
                out.append((mtype, data, msize))
                if mtype == 0x0010:       # continuation
                    blocks.append((self._off(data), self._len(data + self.so)))
                pos = data + msize
        return out

    # ---- groups ---------------------------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        assert self.buf[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = self._off(heap_addr + 8 + 2 * self.sl)
        start = data_addr + off
        end = self.buf.index(b"\x00", start)
        return self.buf[start:end].decode()

    def _btree_entries(self, btree_addr, heap_addr):
        assert self.buf[btree_addr:btree_addr + 4] == b"TREE", "expected a v1 B-tree node"
        node_type, level = self.buf[btree_addr + 4], self.buf[btree_addr + 5]
        n = struct.unpack_from("<H", self.buf, btree_addr + 6)[0]
        assert node_type == 0
        pos = btree_addr + 8 + 2 * self.so
        entries = []
        for i in range(n):
            pos += self.sl                     # key i
            child = self._off(pos)
            pos += self.so
            if level > 0:
                entries += self._btree_entries(child, heap_addr)
            else:
                assert self.buf[child:child + 4] == b"SNOD"
                n_sym = struct.unpack_from("<H", self.buf, child + 6)[0]
                epos = child + 8
                for _ in range(n_sym):
                    e = self._symbol_entry(epos)
                    e["name"] = self._heap_name(heap_addr, e["name_off"])
                    entries.append(e)
                    epos += 2 * self.so + 24
        return entries

    def _group_children(self, header_addr):
        for mtype, data, _ in self._messages(header_addr):
            if mtype == 0x0011:                # symbol table message
                return self._btree_entries(self._off(data), self._off(data + self.so))
        return None

    # ---- datasets -------------------------------------------------------------------------------------------------------
    def _dataset(self, header_addr):
        shape = dtype = address = None
        for mtype, data, _ in self._messages(header_addr):
            if mtype == 0x0001:                # dataspace
                version, rank = self.buf[data], self.buf[data + 1]
                dims_pos = data + (8 if version == 1 else 4)
                shape = tuple(self._len(dims_pos + i * self.sl) for i in range(rank))
            elif mtype == 0x0003:              # datatype
                cls = self.buf[data] & 0x0F
                size = struct.unpack_from("<I", self.buf, data + 4)[0]
                if cls != 1:
                    return None                # not floating point
                dtype = {4: "<f4", 8: "<f8", 2: "<f2"}[size]
            elif mtype == 0x0008:              # layout
                version = self.buf[data]
                if version == 3:
                    if self.buf[data + 1] != 1:
                        raise ValueError("only contiguous datasets are supported")
                    address = self._off(data + 2)
                else:
                    raise ValueError(f"layout message version {version} not supported")
        if shape is None or dtype is None or address is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=self.base + address)
        return arr.reshape(shape).copy()

    # ---- public -----------------------------------------------------------------------------------------------------------
    def datasets(self):
        """{'/group/.../name': ndarray} for every float dataset in the file."""
        out = {}

        def walk(header_addr, prefix):
            children = self._group_children(header_addr)
            if children is None:
                arr = self._dataset(header_addr)
                if arr is not None:
                    out[prefix] = arr
                return
            for e in children:
                walk(e["header"], prefix + "/" + e["name"])
        walk(self.root["header"], "")
        return out


def load_keras_nerf_weights(path):
    """Flat fp32 parameter vectors (coarse, fine) in the order [W0, b0, W1, b1, ...] of Keras layer creation.

    The reference's NeRF(keras.Model) holds two functional models: `model` (coarse: dense ... dense_10) and `model_1`
    (fine: dense_11 ... dense_21) -- SURVEY.md section 4."""
    ds = H5File(path).datasets()
    layers = {}
    for key, arr in ds.items():
        parts = key.strip("/").split("/")
        name = parts[-1]                      # kernel:0 / bias:0
        layer = parts[-2]                     # dense_7
        idx = 0 if layer == "dense" else int(layer.split("_")[1])
        layers.setdefault(idx, {})[name.split(":")[0]] = arr.astype(np.float32)
    order = sorted(layers)
    half = len(order) // 2

    def flat(ids):
        return np.concatenate([np.concatenate([layers[i]["kernel"].reshape(-1), layers[i]["bias"].reshape(-1)])
                               for i in ids])
    return flat(order[:half]), flat(order[half:]), {i: layers[i]["kernel"].shape for i in order}
