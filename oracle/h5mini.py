"""Minimal read-only HDF5 reader for the reference's fixture file -- TEST INFRASTRUCTURE (container only).

``h5py`` is not installed and there is no network, but the reference's ``tests/test_data.h5`` (7 converged orbits,
reference ``tests/test_basic.py:346-387``) uses only old-style structures: superblock v0, v1 object headers,
symbol-table groups (B-tree v1 + local heap), contiguous little-endian f64 datasets, fixed-size and variable-length
string attributes (global heap).  That subset is parsed here so that ``oracle/make_golden.py``
can lift the stored orbits into ``tests/golden/``.  Not a general HDF5 implementation.
"""
import struct

import numpy as np


class H5Mini:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.d = fh.read()
        d = self.d
        assert d[:8] == b"\x89HDF\r\n\x1a\n" and d[8] == 0, "superblock v0 expected"
        assert d[13] == 8 and d[14] == 8
        # root symbol-table entry at 56: link-name offset, object header address, cache type, reserved, scratch
        _, self.root_header, cache_type, _ = struct.unpack_from("<QQII", d, 56)
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", d, 56 + 24)

    # -- groups ---------------------------------------------------------------------------------------------------
    def _heap_data(self, heap_addr):
        assert self.d[heap_addr : heap_addr + 4] == b"HEAP"
        return struct.unpack_from("<Q", self.d, heap_addr + 24)[0]

    def _name(self, heap_data, off):
        end = self.d.index(b"\x00", heap_data + off)
        return self.d[heap_data + off : end].decode()

    def _walk_btree(self, addr, heap_data, out):
        d = self.d
        assert d[addr : addr + 4] == b"TREE"
        level, used = d[addr + 5], struct.unpack_from("<H", d, addr + 6)[0]
        pos = addr + 24
        for _ in range(used):
            child = struct.unpack_from("<Q", d, pos + 8)[0]  # key (8) then child (8)
            pos += 16
            if level > 0:
                self._walk_btree(child, heap_data, out)
            else:
                assert d[child : child + 4] == b"SNOD"
                count = struct.unpack_from("<H", d, child + 6)[0]
                for i in range(count):
                    e = child + 8 + 40 * i
                    name_off, header = struct.unpack_from("<QQ", d, e)
                    out[self._name(heap_data, name_off)] = header

    def _group_entries(self, btree, heap):
        out = {}
        self._walk_btree(btree, self._heap_data(heap), out)
        return out

    # -- object headers -------------------------------------------------------------------------------------------
    def _messages(self, addr):
        d = self.d
        assert d[addr] == 1, "v1 object header expected"
        nmsg = struct.unpack_from("<H", d, addr + 2)[0]
        size = struct.unpack_from("<I", d, addr + 8)[0]
        blocks = [(addr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize = struct.unpack_from("<HH", d, pos)
                body = pos + 8
                if mtype == 0x10:
                    blocks.append(struct.unpack_from("<QQ", d, body))
                msgs.append((mtype, body, msize))
                pos = body + msize
        return msgs

    @staticmethod
    def _pad8(x):
        return (x + 7) & ~7

    def _parse_datatype(self, pos):
        d = self.d
        cls = d[pos] & 0x0F
        size = struct.unpack_from("<I", d, pos + 4)[0]
        base = None
        if cls == 9:  # variable length: base type follows the 8-byte header
            base = self._parse_datatype(pos + 8)
        return {"class": cls, "size": size, "bits0": d[pos + 1], "base": base}

    def _parse_dataspace(self, pos):
        d = self.d
        ver, rank = d[pos], d[pos + 1]
        assert ver == 1
        return tuple(struct.unpack_from("<" + "Q" * rank, d, pos + 8)) if rank else ()

    def _global_heap_object(self, addr, index):
        d = self.d
        assert d[addr : addr + 4] == b"GCOL"
        total = struct.unpack_from("<Q", d, addr + 8)[0]
        pos = addr + 16
        while pos < addr + total:
            idx, _, _, size = struct.unpack_from("<HHIQ", d, pos)
            if idx == index:
                return d[pos + 16 : pos + 16 + size]
            if idx == 0:
                break
            pos += 16 + self._pad8(size)
        raise KeyError(index)

    def _decode(self, dtype, shape, pos):
        d = self.d
        count = int(np.prod(shape)) if shape else 1
        if dtype["class"] == 1:
            arr = np.frombuffer(d, dtype="<f8" if dtype["size"] == 8 else "<f4", count=count, offset=pos)
            return arr.reshape(shape).copy() if shape else float(arr[0])
        if dtype["class"] == 0:
            signed = bool(dtype["bits0"] & 0x08)
            code = {1: "b", 2: "h", 4: "i", 8: "q"}[dtype["size"]]
            arr = np.frombuffer(d, dtype="<" + (code if signed else code.upper()), count=count, offset=pos)
            return arr.reshape(shape).copy() if shape else int(arr[0])
        if dtype["class"] == 3:
            return d[pos : pos + dtype["size"]].split(b"\x00")[0].decode()
        if dtype["class"] == 9:
            length, gaddr, gidx = struct.unpack_from("<IQI", d, pos)
            return self._global_heap_object(gaddr, gidx)[:length].decode()
        raise NotImplementedError(dtype)

    def _attribute(self, body):
        d = self.d
        ver = d[body]
        assert ver == 1
        name_sz, dt_sz, ds_sz = struct.unpack_from("<HHH", d, body + 2)
        pos = body + 8
        name = d[pos : pos + name_sz].split(b"\x00")[0].decode()
        pos += self._pad8(name_sz)
        dtype = self._parse_datatype(pos)
        pos += self._pad8(dt_sz)
        shape = self._parse_dataspace(pos)
        pos += self._pad8(ds_sz)
        return name, self._decode(dtype, shape, pos)

    def read_object(self, addr):
        """Returns ("group", {name: header_addr}) or ("dataset", ndarray, attrs)."""
        msgs = self._messages(addr)
        kinds = {m[0] for m in msgs}
        if 0x11 in kinds:
            body = [m for m in msgs if m[0] == 0x11][0][1]
            btree, heap = struct.unpack_from("<QQ", self.d, body)
            return "group", self._group_entries(btree, heap)
        dtype = shape = data_addr = None
        attrs = {}
        for mtype, body, _ in msgs:
            if mtype == 0x01:
                shape = self._parse_dataspace(body)
            elif mtype == 0x03:
                dtype = self._parse_datatype(body)
            elif mtype == 0x08:
                assert self.d[body] == 3 and self.d[body + 1] == 1, "contiguous layout v3 expected"
                data_addr = struct.unpack_from("<Q", self.d, body + 2)[0]
            elif mtype == 0x0C:
                k, v = self._attribute(body)
                attrs[k] = v
        return "dataset", self._decode(dtype, shape, data_addr), attrs

    def datasets(self):
        """{"group/dataset": (array, attrs)} for the two-level layout of test_data.h5."""
        out = {}
        for gname, gaddr in self._group_entries(self.root_btree, self.root_heap).items():
            kind, entries = self.read_object(gaddr)[:2]
            if kind != "group":
                continue
            for dname, daddr in entries.items():
                obj = self.read_object(daddr)
                if obj[0] == "dataset":
                    out[f"{gname}/{dname}"] = (obj[1], obj[2])
        return out
