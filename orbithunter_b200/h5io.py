"""On-disk format of the reference: Orbit.to_h5 / read_h5 (orbithunter/core.py:1855-1950, orbithunter/io.py:25-139) without h5py.

The reference stores one HDF5 dataset per orbit (the state array) with the attributes ``basis``, ``class``, ``discretization``,
``parameters`` (+ ``frame`` for the relative classes, ``cost`` on request), datasets grouped by name paths.  This image has no
h5py / libhdf5, so the subset of HDF5 the reference's own files use (its tests/test_data.h5 was written by h5py) is written and
read here directly: superblock version 0, version-1 object headers, symbol-table groups (version-1 B-tree + local heap),
contiguous little-endian float64 datasets, float64 / int64 / variable-length UTF-8 string attributes (global heap).
Files are rewritten as a whole on every save (append = read, merge, write), which keeps the writer a single forward pass.

``to_h5_batch`` is the batch writer SURVEY 8(f)-4 asks for: every orbit of a BatchedOrbitKS in one pass, costs evaluated by one
device call.
"""
from __future__ import annotations

import os
import struct
import time

import numpy as np

__all__ = ["to_h5", "to_h5_batch", "read_h5", "h5_tree", "read_tileset", "to_symbol_string", "to_symbol_array"]

_UNDEF = 0xFFFFFFFFFFFFFFFF
_SIGNATURE = b"\x89HDF\r\n\x1a\n"
# datatype messages, byte for byte as libhdf5 encodes them (cf. the reference fixture tests/test_data.h5)
_DT_F64 = bytes.fromhex("11203f000800000000004000340b0034ff030000")
_DT_I64 = bytes.fromhex("100800000800000000004000")
_DT_VSTR = bytes.fromhex("1901010010000000" "1000000001000000" "00000800")  # vlen of 1-byte chars, string, UTF-8


def _pad8(n):
    return (n + 7) & ~7


# ---------------------------------------------------------------- writer -----------------------------------------------------
class _Writer:
    """Lays the file out front to back: every structure is built as bytes once its children's addresses are known."""

    def __init__(self, leaf_k):
        self.buf = bytearray(96)  # superblock, filled in last
        self.leaf_k = leaf_k
        self.strings = []  # global-heap objects of the collection being filled
        self.gcol_fixups = []  # (position in buf of a vlen descriptor's address field)

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- messages
    @staticmethod
    def _msg(mtype, body, flags=0):
        body = body + b"\x00" * (_pad8(len(body)) - len(body))
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    @staticmethod
    def _dataspace(shape):
        if not shape:
            return struct.pack("<BBB5x", 1, 0, 0)
        return struct.pack("<BBB5x", 1, len(shape), 1) + b"".join(struct.pack("<Q", s) for s in shape) * 2

    def _vlen(self, text):
        raw = text.encode("utf-8")
        self.strings.append(raw)
        return raw, len(self.strings)  # (bytes, 1-based index in the collection)

    def _attribute(self, name, value):
        nm = name.encode() + b"\x00"
        if isinstance(value, str):
            dt, ds = _DT_VSTR, self._dataspace(())
            raw, idx = self._vlen(value)
            data = struct.pack("<IQI", len(raw), 0, idx)  # collection address patched when the collection is placed
            fix = True
        else:
            arr = np.asarray(value)
            if arr.dtype.kind in "iub":
                dt, arr = _DT_I64, arr.astype("<i8")
            else:
                dt, arr = _DT_F64, arr.astype("<f8")
            ds, data, fix = self._dataspace(arr.shape), arr.tobytes(), False
        head = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds))
        body = head + nm.ljust(_pad8(len(nm)), b"\x00") + dt.ljust(_pad8(len(dt)), b"\x00") + ds.ljust(_pad8(len(ds)), b"\x00")
        return body + data, (len(body) + 4 if fix else None)  # offset of the address field inside the message body

    def _object_header(self, messages):
        """messages: list of (bytes, fixup offset inside the message or None) -> address; version-1 header, one chunk."""
        blob = b"".join(m for m, _ in messages)
        head = struct.pack("<BxHII4x", 1, len(messages), 1, len(blob))
        addr = self.alloc(head + blob)
        pos = addr + 16
        for m, fix in messages:
            if fix is not None:
                self.gcol_fixups.append(pos + 8 + fix)
            pos += len(m)
        return addr

    def dataset(self, array, attrs):
        arr = np.ascontiguousarray(array, dtype="<f8")
        data_addr = self.alloc(arr.tobytes())
        msgs = [(self._msg(0x01, self._dataspace(arr.shape)), None), (self._msg(0x03, _DT_F64, 1), None),
                (self._msg(0x05, bytes.fromhex("0202020100000000"), 1), None),
                (self._msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)), None),
                (self._msg(0x12, struct.pack("<B3xI", 1, int(time.time()) & 0xFFFFFFFF)), None)]
        for k, v in attrs.items():
            body, fix = self._attribute(k, v)
            msgs.append((self._msg(0x0C, body, 4), fix))
        return self._object_header(msgs)

    def group(self, entries):
        """entries: {name: (object header address, (btree, heap) for groups or None)} -> (header, btree, heap) addresses."""
        names = sorted(entries, key=lambda s: s.encode())
        seg, offsets = bytearray(8), {}
        for nm in names:
            offsets[nm] = len(seg)
            raw = nm.encode() + b"\x00"
            seg += raw.ljust(_pad8(len(raw)), b"\x00")
        free_off = len(seg)
        seg += struct.pack("<QQ", 1, 16)  # one minimal free block terminates the free list (H5HL_FREE_NULL = 1)
        heap_data = self.alloc(bytes(seg))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, heap_data))
        cap = 2 * self.leaf_k
        assert len(names) <= cap
        snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(names)))
        for nm in names:
            hdr, sub = entries[nm]
            if sub is None:
                snod += struct.pack("<QQII16x", offsets[nm], hdr, 0, 0)
            else:
                snod += struct.pack("<QQIIQQ", offsets[nm], hdr, 1, 0, sub[0], sub[1])
        snod += b"\x00" * (8 + cap * 40 - len(snod))
        ik = 16
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, _UNDEF, _UNDEF))
        if names:
            snod_addr = self.alloc(bytes(snod))
            tree += struct.pack("<QQQ", 0, snod_addr, offsets[names[-1]])
        tree += b"\x00" * (24 + (2 * ik + 1) * 8 + 2 * ik * 8 - len(tree))
        btree = self.alloc(bytes(tree))
        header = self._object_header([(self._msg(0x11, struct.pack("<QQ", btree, heap)), None)])
        return header, btree, heap

    def flush_strings(self):
        """Place the global heap collection holding every vlen string written so far and patch the descriptors."""
        if not self.strings:
            return
        body = bytearray()
        for i, raw in enumerate(self.strings, start=1):
            body += struct.pack("<HHIQ", i, 0, 0, len(raw)) + raw.ljust(_pad8(len(raw)), b"\x00")
        size = max(4096, _pad8(16 + len(body) + 16))
        free = size - 16 - len(body)
        body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\x00" * (free - 16)
        addr = self.alloc(b"GCOL" + struct.pack("<B3xQ", 1, size) + bytes(body))
        for pos in self.gcol_fixups:
            struct.pack_into("<Q", self.buf, pos, addr)
        self.strings, self.gcol_fixups = [], []

    def finish(self, root):
        header, btree, heap = root
        self.flush_strings()
        while len(self.buf) % 8:
            self.buf.append(0)
        sb = _SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, 16, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQIIQQ", 0, header, 1, 0, btree, heap)
        self.buf[:96] = sb
        return bytes(self.buf)


def _largest_group(tree):
    size = len(tree)
    for v in tree.values():
        if isinstance(v, dict):
            size = max(size, _largest_group(v))
    return size


def _write_tree(w, tree):
    entries = {}
    for name, node in tree.items():
        if isinstance(node, dict):
            hdr, bt, hp = _write_tree(w, node)
            entries[name] = (hdr, (bt, hp))
        else:
            array, attrs = node
            entries[name] = (w.dataset(array, attrs), None)
            if len(w.strings) > 60000:  # global-heap object indices are 16 bit
                w.flush_strings()
    return w.group(entries)


def write_tree(filename, tree):
    """tree: nested dict, leaves (ndarray, {attribute: str | int/float scalar or array}).  Overwrites ``filename``."""
    w = _Writer(max(4, (_largest_group(tree) + 1) // 2))
    data = w.finish(_write_tree(w, tree))
    tmp = filename + ".tmp"
    with open(tmp, "wb") as fh:
        fh.write(data)
    os.replace(tmp, filename)


# ---------------------------------------------------------------- reader -----------------------------------------------------
class _Reader:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.d = fh.read()
        d = self.d
        if d[:8] != _SIGNATURE or d[8] != 0 or d[13] != 8 or d[14] != 8:
            raise ValueError(f"{path}: not an HDF5 file of the supported kind (superblock version 0, 8-byte offsets)")
        self.root = struct.unpack_from("<QQ", d, 56 + 24)

    def _cstr(self, pos):
        return self.d[pos:self.d.index(b"\x00", pos)].decode()

    def _entries(self, btree, heap):
        d, out = self.d, {}
        heap_data = struct.unpack_from("<Q", d, heap + 24)[0]

        def walk(node):
            level, used = d[node + 5], struct.unpack_from("<H", d, node + 6)[0]
            for i in range(used):
                child = struct.unpack_from("<Q", d, node + 24 + 16 * i + 8)[0]
                if level:
                    walk(child)
                else:
                    for s in range(struct.unpack_from("<H", d, child + 6)[0]):
                        off, hdr = struct.unpack_from("<QQ", d, child + 8 + 40 * s)
                        out[self._cstr(heap_data + off)] = hdr

        walk(btree)
        return out

    def _messages(self, addr):
        d = self.d
        if d[addr] != 1:
            raise ValueError("version-1 object headers only")
        count, size = struct.unpack_from("<H", d, addr + 2)[0], struct.unpack_from("<I", d, addr + 8)[0]
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < count:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(out) < count:
                mtype, msize = struct.unpack_from("<HH", d, pos)
                if mtype == 0x10:
                    blocks.append(struct.unpack_from("<QQ", d, pos + 8))
                out.append((mtype, pos + 8))
                pos += 8 + msize
        return out

    def _value(self, dt, shape, pos):
        d = self.d
        cls, size = d[dt] & 0x0F, struct.unpack_from("<I", d, dt + 4)[0]
        count = int(np.prod(shape)) if shape else 1
        if cls == 1:
            arr = np.frombuffer(d, dtype=f"<f{size}", count=count, offset=pos)
            return arr.reshape(shape).astype(float) if shape else float(arr[0])
        if cls == 0:
            code = ("i" if d[dt + 1] & 0x08 else "u") + str(size)
            arr = np.frombuffer(d, dtype="<" + code, count=count, offset=pos)
            return arr.reshape(shape).astype(np.int64) if shape else int(arr[0])
        if cls == 3:
            return d[pos:pos + size].split(b"\x00")[0].decode()
        if cls == 9:
            length, gaddr, gidx = struct.unpack_from("<IQI", d, pos)
            total, p = struct.unpack_from("<Q", d, gaddr + 8)[0], gaddr + 16
            while p < gaddr + total:
                idx, _, _, osz = struct.unpack_from("<HHIQ", d, p)
                if idx == gidx:
                    return d[p + 16:p + 16 + length].decode()
                if idx == 0:
                    break
                p += 16 + _pad8(osz)
            raise ValueError("dangling variable-length string")
        raise ValueError(f"unsupported HDF5 datatype class {cls}")

    def _shape(self, pos):
        rank = self.d[pos + 1]
        return tuple(struct.unpack_from("<" + "Q" * rank, self.d, pos + 8)) if rank else ()

    def node(self, addr):
        """-> nested dict for a group, (array, attrs) for a dataset."""
        msgs = self._messages(addr)
        for mtype, body in msgs:
            if mtype == 0x11:
                bt, hp = struct.unpack_from("<QQ", self.d, body)
                return {k: self.node(v) for k, v in self._entries(bt, hp).items()}
        dt = shape = data = None
        attrs = {}
        for mtype, body in msgs:
            if mtype == 0x01:
                shape = self._shape(body)
            elif mtype == 0x03:
                dt = body
            elif mtype == 0x08:
                if self.d[body] != 3 or self.d[body + 1] != 1:
                    raise ValueError("contiguous dataset layout (version 3) only")
                data = struct.unpack_from("<Q", self.d, body + 2)[0]
            elif mtype == 0x0C:
                nsz, dsz, ssz = struct.unpack_from("<HHH", self.d, body + 2)
                p = body + 8
                name = self.d[p:p + nsz].split(b"\x00")[0].decode()
                adt = p + _pad8(nsz)
                asp = adt + _pad8(dsz)
                attrs[name] = self._value(adt, self._shape(asp), asp + _pad8(ssz))
        return self._value(dt, shape, data), attrs

    def tree(self):
        return {k: self.node(v) for k, v in self._entries(*self.root).items()}


def h5_tree(filename):
    """The whole file as a nested dict {name: subtree | (state array, attribute dict)}."""
    return _Reader(filename).tree()


# ---------------------------------------------------------------- orbit level ------------------------------------------------
def _orbit_attrs(orbit, cost=None):
    attrs = {"basis": orbit.basis, "class": orbit.__class__.__name__, "discretization": np.asarray(orbit.discretization, dtype=np.int64),
             "parameters": np.asarray(orbit.parameters, dtype=float)}
    if hasattr(orbit, "frame"):
        attrs["frame"] = orbit.frame
    if cost is not None:
        attrs["cost"] = float(cost)
    return attrs


def _split(groupname, dataname):
    return [p for p in "/".join(groupname.split("/") + dataname.split("/")).split("/") if p]


def _lookup(tree, parts):
    node = tree
    for p in parts:
        if not isinstance(node, dict) or p not in node:
            return None
        node = node[p]
    return node


def _free_name(tree, groupname, dataname):
    """The reference's collision rule (core.py:1902-1918): numeric names count up, other names get _0, _1, ... appended."""
    i = 0
    dataname = dataname or str(i)
    parts = _split(groupname, dataname)
    while _lookup(tree, parts) is not None:
        try:
            dataname = str(int(dataname) + 1)
            parts = _split(groupname, dataname)
        except ValueError:
            parts = _split(groupname, "".join([dataname.rstrip("/"), "_", str(i)]))
            i += 1
    return parts


def _insert(tree, parts, leaf):
    node = tree
    for p in parts[:-1]:
        node = node.setdefault(p, {})
        if not isinstance(node, dict):
            raise ValueError(f"'{p}' is a dataset, not a group")
    node[parts[-1]] = leaf


def _open_tree(filename, h5mode):
    if h5mode not in ("a", "r+", "w", "w-"):
        raise ValueError("h5mode must be one of 'a', 'r+', 'w-', 'w'")
    exists = os.path.exists(filename)
    if h5mode == "w-" and exists:
        raise FileExistsError(filename)
    if h5mode == "r+" and not exists:
        raise FileNotFoundError(filename)
    return h5_tree(filename) if exists and h5mode in ("a", "r+") else {}


def to_h5(orbit, filename="", groupname="", dataname="", h5mode="a", verbose=False, include_cost=False, **kwargs):
    """Orbit.to_h5 (core.py:1855-1950): one dataset (the state) with the orbit's attributes, under groupname/dataname."""
    filename = filename or orbit.filename(extension=".h5")
    tree = _open_tree(filename, h5mode)
    parts = _free_name(tree, groupname, dataname)
    if verbose:
        print('Writing dataset "{}" to file {}'.format("/".join(parts), filename))
    cost = None
    if include_cost:
        try:
            cost = orbit.cost()
        except (ZeroDivisionError, ValueError, AttributeError):
            print(f"Unable to compute cost for instance {repr(orbit)}; data will not be saved to .h5 file.")
    _insert(tree, parts, (np.asarray(orbit.state, dtype=float), _orbit_attrs(orbit, cost)))
    write_tree(filename, tree)
    return "/".join(parts)


def to_h5_batch(batch, filename, groupname="", datanames=None, h5mode="a", include_cost=True, select=None):
    """Every orbit of a BatchedOrbitKS (or those of the index array ``select``) as one dataset each, in ONE write pass; the
    costs come from one device evaluation.  ``datanames`` defaults to the reference's numeric naming ('0', '1', ...)."""
    from .ks import CLASSES

    idx = np.arange(batch.batch) if select is None else np.asarray(select).ravel()
    cost = batch.transform("modes").cost().cpu().numpy() if include_cost else None
    states, params = batch.cpu_state(), batch.cpu_parameters()
    tree = _open_tree(filename, h5mode)
    cls = CLASSES[batch.cls_name]
    written = []
    for pos, b in enumerate(idx):
        o = cls(state=states[b], basis=batch.basis, parameters=tuple(float(x) for x in params[b]), discretization=batch.discretization)
        parts = _free_name(tree, groupname, "" if datanames is None else str(datanames[pos]))
        _insert(tree, parts, (states[b], _orbit_attrs(o, None if cost is None else cost[b])))
        written.append("/".join(parts))
    write_tree(filename, tree)
    return written


def read_h5(filename, *datanames, validate=False, **orbitkwargs):
    """io.read_h5 (io.py:25-139): orbits by dataset / group name (all of the file when no name is given).  A dataset name
    gives an orbit, a group name the list of every orbit below it; several names give a list in that order."""
    from .ks import CLASSES

    if len(datanames) == 1 and isinstance(datanames[0], tuple):
        datanames = tuple(datanames[0])
    tree = h5_tree(os.path.abspath(filename))

    def build(leaf):
        state, attrs = leaf
        if attrs.get("class") not in CLASSES:
            raise NotImplementedError(f"class {attrs.get('class')} is outside this package (OrbitKS, RelativeOrbitKS, "
                                      "ShiftReflectionOrbitKS, AntisymmetricOrbitKS)")
        kw = {k: v for k, v in attrs.items() if k not in ("class", "parameters", "discretization", "cost", "residual")}
        par = attrs.get("parameters")
        disc = attrs.get("discretization")
        kw.update(parameters=None if par is None else tuple(float(x) for x in np.atleast_1d(par)),
                  discretization=None if disc is None else tuple(int(x) for x in np.atleast_1d(disc)))
        kw.update(orbitkwargs)
        return CLASSES[attrs["class"]](state=state, **kw)

    def collect(node, out):
        for v in node.values():
            if isinstance(v, dict):
                collect(v, out)
            else:
                out.append(build(v))
        return out

    imported = []
    for name in datanames or tree:
        node = _lookup(tree, [p for p in str(name).split("/") if p])
        if node is None:
            raise KeyError(name)
        if isinstance(node, dict):
            group = collect(node, [])
            imported.append(group[0] if len(group) == 1 else group)
        else:
            imported.append(build(node))
    if validate and len(imported) == 1:
        return imported[0].preprocess()
    if len(imported) == 1:
        return imported[0]
    if validate:
        return [x.preprocess() for x in imported]
    return imported


def read_tileset(filename, keys, orbit_names, validate=False, **orbitkwargs):
    """io.read_tileset (io.py:142-169): the tiling dictionary {key: orbit} for ``gluing.tile`` from datasets of one file."""
    if len(keys) != len(orbit_names):
        raise AssertionError("as many keys as orbit names are needed")
    orbits = read_h5(filename, tuple(orbit_names), validate=validate, **orbitkwargs)
    return dict(zip(keys, orbits if isinstance(orbits, list) else [orbits]))


def to_symbol_string(symbol_array):
    """io.to_symbol_string (io.py:172-181): the symbols of an array as one string -- [[0, 1], [2, 0]] -> "01_20".  Pass k
    (k = 0, 1, ...) groups the current pieces in runs of ``shape[1 + k]`` and joins each run with k underscores; what is left
    is joined with ``ndim - 1`` underscores.  (Same grouping as the reference, including for arrays whose trailing axes have
    different lengths.)"""
    arr = np.asarray(symbol_array)
    pieces = [str(v) for v in arr.ravel().tolist()]
    for k, run in enumerate(arr.shape[1:]):
        pieces = [("_" * k).join(pieces[p:p + run]) for p in range(0, len(pieces), run)]
    return ("_" * (arr.ndim - 1)).join(pieces)


def to_symbol_array(symbol_string, symbol_array_shape):
    """io.to_symbol_array (io.py:184-189): inverse of ``to_symbol_string`` for single-digit symbols."""
    return np.array([ch for ch in symbol_string.replace("_", "")]).astype(int).reshape(symbol_array_shape)
