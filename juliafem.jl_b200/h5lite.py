"""Minimal read-only HDF5 parser (pure Python + numpy), just enough for Code Aster / Salome `.med` meshes.

The reference reads `.med` files with `h5read(fn, "/")` (src/readers/read_aster_mesh.jl:38-40) and then picks
`ENS_MAA/<mesh>/<increment>/NOE/{COO,FAM,NUM}`, `.../MAI/<type>/{NOD,FAM,NUM}` and `FAS/<mesh>/{NOEUD,ELEME}/*/GRO/NOM`
(:69-150).  `h5py` is not available in this image, so the subset of the file format those files use is parsed here:

  * superblock version 2 / 3 (root object header address) and version 0 / 1 (root symbol-table entry);
  * version-2 object headers (`OHDR` + `OCHK` continuation chunks) and version-1 object headers;
  * groups stored as compact link messages (0x06), or old-style symbol tables (B-tree v1 + local heap + `SNOD`);
  * datasets: dataspace v1/v2, datatypes fixed-point / IEEE float / fixed-length string, data layout v3 compact or
    contiguous (and v4 contiguous).  Chunked, filtered or dense-link (fractal-heap) storage raises `H5Unsupported`.

`read(path)` returns nested dicts of numpy arrays, the shape `h5read(fn, "/")` gives the reference (Julia's HDF5 reverses
the dimension order; callers reshape explicitly, see mesh.read_med).
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Unsupported(NotImplementedError):
    pass


class _File:
    def __init__(self, data: bytes):
        self.d = data
        if data[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file (bad signature)")
        ver = data[8]
        if ver in (2, 3):
            so, sl = data[9], data[10]
            if (so, sl) != (8, 8):
                raise H5Unsupported("only 8-byte offsets/lengths are supported")
            self.base, _ext, _eof, root = struct.unpack_from("<QQQQ", data, 12)
            self.root = ("ohdr", root)
        elif ver in (0, 1):
            so, sl = data[13], data[14]
            if (so, sl) != (8, 8):
                raise H5Unsupported("only 8-byte offsets/lengths are supported")
            p = 24 if ver == 0 else 28
            self.base = struct.unpack_from("<Q", data, p)[0]
            p += 32  # base, free-space, eof, driver
            # root symbol table entry: link name offset, object header address, cache type, reserved, scratch
            _lno, oh, cache = struct.unpack_from("<QQI", data, p)
            self.root = ("ohdr", oh)
        else:
            raise H5Unsupported(f"superblock version {ver}")

    # ---- object headers -------------------------------------------------------------------------------------------
    def messages(self, off: int):
        d = self.d
        off += self.base
        if d[off:off + 4] == b"OHDR":
            return self._messages_v2(off)
        return self._messages_v1(off)

    def _messages_v2(self, off: int):
        d = self.d
        flags = d[off + 5]
        p = off + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szb = 1 << (flags & 3)
        chunk0 = int.from_bytes(d[p:p + szb], "little")
        p += szb
        blocks = [(p, p + chunk0)]
        out = []
        while blocks:
            s, e = blocks.pop(0)
            q = s
            while q + 4 <= e:
                t = d[q]
                sz = struct.unpack_from("<H", d, q + 1)[0]
                q += 4
                if flags & 0x04:
                    q += 2
                body = d[q:q + sz]
                q += sz
                if t == 0x10:
                    co, cl = struct.unpack_from("<QQ", body, 0)
                    co += self.base
                    if d[co:co + 4] != b"OCHK":
                        raise ValueError("bad object-header continuation chunk")
                    blocks.append((co + 4, co + cl - 4))
                else:
                    out.append((t, body))
        return out

    def _messages_v1(self, off: int):
        d = self.d
        if d[off] != 1:
            raise ValueError(f"unknown object header at {off:#x}")
        nmsg = struct.unpack_from("<H", d, off + 2)[0]
        hsize = struct.unpack_from("<I", d, off + 8)[0]
        blocks = [(off + 16, off + 16 + hsize)]
        out = []
        while blocks and len(out) < nmsg + 64:
            s, e = blocks.pop(0)
            q = s
            while q + 8 <= e:
                t, sz = struct.unpack_from("<HH", d, q)
                q += 8
                body = d[q:q + sz]
                q += sz
                if t == 0x10:
                    co, cl = struct.unpack_from("<QQ", body, 0)
                    blocks.append((co + self.base, co + self.base + cl))
                else:
                    out.append((t, body))
        return out

    # ---- groups ---------------------------------------------------------------------------------------------------
    def links(self, msgs):
        """name -> object header address, from link messages or an old-style symbol table."""
        out = {}
        for t, b in msgs:
            if t == 0x06:
                f = b[1]
                p = 2
                ltype = 0
                if f & 0x08:
                    ltype = b[p]
                    p += 1
                if f & 0x04:
                    p += 8
                if f & 0x10:
                    p += 1
                ls = 1 << (f & 3)
                ln = int.from_bytes(b[p:p + ls], "little")
                p += ls
                name = b[p:p + ln].decode()
                p += ln
                if ltype == 0:
                    out[name] = struct.unpack_from("<Q", b, p)[0]
            elif t == 0x02:
                # link info: fractal heap address != undefined means dense link storage
                f = b[1]
                p = 2 + (8 if f & 1 else 0)
                heap = struct.unpack_from("<Q", b, p)[0]
                if heap != UNDEF:
                    raise H5Unsupported("dense (fractal-heap) group storage")
            elif t == 0x11:
                btree, heap = struct.unpack_from("<QQ", b, 0)
                out.update(self._symbol_table(btree, heap))
        return out

    def _symbol_table(self, btree: int, heap: int):
        d = self.d
        hp = heap + self.base
        if d[hp:hp + 4] != b"HEAP":
            raise ValueError("bad local heap")
        hdata = struct.unpack_from("<Q", d, hp + 24)[0] + self.base
        out = {}

        def name_at(o):
            e = d.index(b"\0", hdata + o)
            return d[hdata + o:e].decode()

        def node(addr):
            a = addr + self.base
            if d[a:a + 4] == b"TREE":
                level, used = d[a + 5], struct.unpack_from("<H", d, a + 6)[0]
                p = a + 24
                for i in range(used):
                    p += 8  # key
                    child = struct.unpack_from("<Q", d, p)[0]
                    p += 8
                    node(child)
            elif d[a:a + 4] == b"SNOD":
                n = struct.unpack_from("<H", d, a + 6)[0]
                p = a + 8
                for i in range(n):
                    lno, oh = struct.unpack_from("<QQ", d, p)
                    out[name_at(lno)] = oh
                    p += 40
            else:
                raise ValueError("bad group B-tree node")

        node(btree)
        return out

    # ---- datasets -------------------------------------------------------------------------------------------------
    @staticmethod
    def _dataspace(b):
        ver, rank, flags = b[0], b[1], b[2]
        p = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", b, p + 8 * i)[0] for i in range(rank))

    @staticmethod
    def _datatype(b):
        cls = b[0] & 0x0F
        bits0 = b[1]
        size = struct.unpack_from("<I", b, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise H5Unsupported(f"datatype class {cls}")

    def dataset(self, msgs):
        shape = dtype = None
        raw = None
        for t, b in msgs:
            if t == 0x01:
                shape = self._dataspace(b)
            elif t == 0x03:
                dtype = self._datatype(b)
            elif t == 0x0B:
                raise H5Unsupported("filtered (compressed) dataset")
            elif t == 0x08:
                ver = b[0]
                if ver not in (3, 4):
                    raise H5Unsupported(f"data layout message version {ver}")
                lc = b[1]
                if lc == 0:
                    n = struct.unpack_from("<H", b, 2)[0]
                    raw = bytes(b[4:4 + n])
                elif lc == 1:
                    addr, n = struct.unpack_from("<QQ", b, 2)
                    raw = b"" if addr == UNDEF else self.d[addr + self.base:addr + self.base + n]
                else:
                    raise H5Unsupported("chunked dataset storage")
        if shape is None or dtype is None or raw is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(raw, dtype=dtype, count=n) if len(raw) >= n * dtype.itemsize else np.zeros(n, dtype)
        return arr.reshape(shape).astype(dtype.newbyteorder("="))

    def tree(self, off: int, depth: int = 0):
        if depth > 32:
            raise ValueError("group nesting too deep")
        msgs = self.messages(off)
        kinds = {t for t, _ in msgs}
        if 0x08 in kinds:
            return self.dataset(msgs)
        return {name: self.tree(a, depth + 1) for name, a in self.links(msgs).items()}


def read(path: str) -> dict:
    """Whole file as nested dicts (groups) of numpy arrays (datasets), like `h5read(path, "/")`."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    return f.tree(f.root[1])


# ---------------------------------------------------------------------------------------------------------------------
# Writer: the classic HDF5 layout (superblock version 0, version-1 object headers, root group as a symbol table = B-tree
# v1 + local heap + one SNOD leaf, contiguous little-endian datasets).  Enough for the heavy data of the XDMF result files
# (xdmf.py), where the reference writes flat `/DataItem_N` datasets with HDF5.jl (src/io.jl:268-312).
# ---------------------------------------------------------------------------------------------------------------------

def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _datatype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 8:
            bits, props = (0x20, 63, 0), struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        else:
            bits, props = (0x20, 31, 0), struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBI", 0x11, bits[0], bits[1], bits[2], dt.itemsize) + props
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise H5Unsupported(f"cannot write dtype {dt}")


def write(path: str, datasets: dict) -> None:
    """Write `datasets` (name -> numpy array; float32/64 and integer types, any rank) as root-level contiguous datasets of a
    new HDF5 file.  The dataspace carries the numpy (C-order) shape, as h5py would."""
    names = sorted(datasets, key=lambda s: s.encode())
    if not names:
        names = []
    for n in names:
        if not n or "/" in n or "\0" in n:
            raise ValueError(f"bad dataset name {n!r} (flat names only)")
    arrays = {}
    for n in names:
        a = np.ascontiguousarray(datasets[n])
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        arrays[n] = a
    leaf_k = max(4, (len(names) + 1) // 2)          # one SNOD leaf holds up to 2 * leaf_k symbols
    internal_k = 16
    if leaf_k > 32767:
        raise H5Unsupported("too many datasets for a single symbol-table leaf")

    # ---- local heap data: the empty string at offset 0, then the names (8-byte aligned)
    heap = bytearray(_pad8(b"\0"))
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode() + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)               # one free block: next = H5HL_FREE_NULL (1), size 16

    # ---- addresses (everything 8-byte aligned, laid out in file order)
    pos = 96                                        # superblock (version 0, 8-byte offsets / lengths)
    root_ohdr = pos
    root_msgs = _msg(0x11, struct.pack("<QQ", 0, 0))    # patched below
    pos += 16 + len(root_msgs)
    heap_hdr = pos
    pos += 32
    heap_data = pos
    pos += len(heap)
    btree = pos
    btree_size = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
    pos += btree_size
    snod = pos
    snod_size = 8 + 2 * leaf_k * 40
    pos += snod_size
    ohdr_addr, data_addr = {}, {}
    hdrs = {}
    for n in names:
        a = arrays[n]
        space = struct.pack("<BBBB4x", 1, a.ndim, 0, 0) + b"".join(struct.pack("<Q", s) for s in a.shape)
        fill = struct.pack("<BBBB", 2, 1, 0, 0)     # version 2, early allocation, fill value undefined
        hdrs[n] = [space, _datatype_msg(a.dtype), fill]
        ohdr_addr[n] = pos
        body_len = sum(len(_msg(t, b)) for t, b in zip((0x01, 0x03, 0x05), hdrs[n])) + len(_msg(0x08, b"\0" * 18))
        pos += 16 + body_len
    for n in names:
        data_addr[n] = pos if arrays[n].nbytes else UNDEF
        pos += arrays[n].nbytes + (-arrays[n].nbytes % 8)
    eof = pos

    out = bytearray()
    # ---- superblock
    out += b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", leaf_k, internal_k, 0)
    out += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    out += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", btree, heap_hdr)     # root symbol-table entry (cached)
    assert len(out) == 96
    # ---- root group object header: one symbol-table message
    root_msgs = _msg(0x11, struct.pack("<QQ", btree, heap_hdr))
    out += struct.pack("<BBHII4x", 1, 0, 1, 1, len(root_msgs)) + root_msgs
    # ---- local heap
    assert len(out) == heap_hdr
    out += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data)
    out += heap
    # ---- B-tree node (group node, level 0, one child)
    assert len(out) == btree
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
    node += struct.pack("<QQQ", 0, snod, name_off[names[-1]] if names else 0)
    out += node + b"\0" * (btree_size - len(node))
    # ---- symbol-table leaf
    assert len(out) == snod
    leaf = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        leaf += struct.pack("<QQII16x", name_off[n], ohdr_addr[n], 0, 0)
    out += leaf + b"\0" * (snod_size - len(leaf))
    # ---- dataset object headers
    for n in names:
        assert len(out) == ohdr_addr[n]
        a = arrays[n]
        layout = struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes)
        msgs = b"".join(_msg(t, b) for t, b in zip((0x01, 0x03, 0x05), hdrs[n])) + _msg(0x08, layout)
        out += struct.pack("<BBHII4x", 1, 0, 4, 1, len(msgs)) + msgs
    # ---- raw data
    for n in names:
        a = arrays[n]
        if a.nbytes:
            assert len(out) == data_addr[n]
            out += a.tobytes() + b"\0" * (-a.nbytes % 8)
    assert len(out) == eof
    with open(path, "wb") as fh:
        fh.write(bytes(out))
