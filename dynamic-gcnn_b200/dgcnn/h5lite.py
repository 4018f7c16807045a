"""dgcnn.h5lite -- a small, dependency-free HDF5 reader / writer for the dense arrays of the reference's `io_h5`
(/root/reference/dgcnn/iotool.py:199-280: `h5py.File(f)[DATA_KEY | LABEL_KEY | WEIGHT_KEY]` in, PyTables EArrays
`DATA_KEY / softmax / LABEL_KEY` (zlib level 5, extendable along axis 0) out).  Neither h5py nor PyTables (nor libhdf5)
exists in this image, so the file format itself is implemented here, following the HDF5 File Format Specification 3.0.

Writer (`write(path, {name: array}, compress=5)`): the classic layout every HDF5 library since 1.0 reads -- superblock
version 0, root group as symbol table (v1 B-tree + local heap + symbol-table node), version-1 object headers with
dataspace / datatype / fill-value / layout (v3) [/ filter pipeline] messages; datasets contiguous, or chunked (one entry
of axis 0 per chunk, unlimited along axis 0) + deflate through a v1 chunk B-tree, plus the CLASS / VERSION / EXTDIM / TITLE
attributes PyTables puts on an EArray so that `tables.open_file` sees what the reference's writer would have produced.

Reader (`File(path)[name]` -> numpy array, `.keys()`, `.attrs(name)`): superblock 0 / 1 (h5py and PyTables defaults) and
2 / 3 (libver='latest'); old-style groups and new-style groups with compact link messages; v1 and v2 object headers with
continuation blocks; fixed-point and IEEE float datatypes of either byte order; compact, contiguous and chunked (v1
B-tree; v4 single-chunk / implicit index) layouts; deflate, shuffle and fletcher32 filters.  Anything else (dense groups,
fixed / extensible-array chunk indexes, compound or variable-length types, external links) raises NotImplementedError
naming the feature.  Nested groups are reached with "a/b/c" paths.

Verification: byte-level known answers taken from the specification, round trips, and the reference's own io_h5 running
on these files through h5py / PyTables stand-ins kept with the test infrastructure (tests/test_h5lite.py).  libhdf5 / h5py / PyTables are absent from this image, so
interoperability with them is by construction from the specification, not by test.

Host-side IO, out of the hot path.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
GROUP_LEAF_K, GROUP_INTERNAL_K, CHUNK_K = 4, 16, 32       # superblock-0 defaults (chunk K is implied there)
HEAP_FREE_NULL = 1                                         # local heap: "no free block" (H5HL_FREE_NULL)

MSG_DATASPACE, MSG_LINK_INFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LINK, MSG_LAYOUT = 0x1, 0x2, 0x3, 0x4, 0x5, 0x6, 0x8
MSG_FILTERS, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE = 0xB, 0xC, 0x10, 0x11


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ====================================================================================================== datatypes
def _encode_datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    big = dt.byteorder == ">"
    if dt.kind == "f" and dt.itemsize in (2, 4, 8):
        exp_bits, man_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        bits0 = (1 if big else 0) | 0x20                       # mantissa normalisation: msb implied
        head = struct.pack("<BBBBI", 0x11, bits0, dt.itemsize * 8 - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        bits0 = (1 if big else 0) | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":                                          # fixed-length, null-terminated ASCII (attributes)
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)
    raise NotImplementedError("h5lite: cannot store dtype %s" % dt)


def _decode_datatype(b: bytes) -> np.dtype:
    cls, bits0, bits1, _bits2, size = struct.unpack_from("<BBBBI", b, 0)
    cls &= 0x0F
    if cls == 0:
        order = ">" if bits0 & 1 else "<"
        return np.dtype("%s%s%d" % (order, "i" if bits0 & 0x08 else "u", size))
    if cls == 1:
        if bits0 & 0x40:
            raise NotImplementedError("h5lite: VAX-endian floats")
        _off, prec, _el, exp_bits, _ml, man_bits, _bias = struct.unpack_from("<HHBBBBI", b, 8)
        if (size, prec, exp_bits, man_bits) not in ((2, 16, 5, 10), (4, 32, 8, 23), (8, 64, 11, 52)):
            raise NotImplementedError("h5lite: non-IEEE float (size %d, exponent %d, mantissa %d)" % (size, exp_bits, man_bits))
        return np.dtype("%sf%d" % (">" if bits0 & 1 else "<", size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise NotImplementedError("h5lite: datatype class %d (only integers, IEEE floats and fixed strings)" % cls)


# ========================================================================================================= writer
class _Out(object):
    """append-only file image; every structure is 8-byte aligned"""

    def __init__(self):
        self.buf = bytearray()

    def tell(self) -> int:
        return len(self.buf)

    def put(self, b: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        at = len(self.buf)
        self.buf += b
        return at


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


def _dataspace(shape, maxshape=None) -> bytes:
    b = struct.pack("<BBBB4x", 1, len(shape), 1 if maxshape is not None else 0, 0)
    b += b"".join(struct.pack("<Q", int(d)) for d in shape)
    if maxshape is not None:
        b += b"".join(struct.pack("<Q", UNDEF if m is None else int(m)) for m in maxshape)
    return b


def _attribute(name: str, value) -> bytes:
    """version-1 attribute message: a fixed string or an integer scalar"""
    if isinstance(value, (bytes, str)):
        raw = value.encode() if isinstance(value, str) else value
        raw = raw + b"\0"
        dt, data = _encode_datatype(np.dtype("S%d" % len(raw))), raw
    else:
        arr = np.asarray(value)
        dt, data = _encode_datatype(arr.dtype), arr.tobytes()
    sp = struct.pack("<BBBB4x", 1, 0, 0, 0)                       # scalar dataspace (rank 0)
    nm = name.encode() + b"\0"
    return struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + data


def _chunk_btree(out: _Out, entries: List[Tuple[Tuple[int, ...], int, int]], rank: int, end_key: Tuple[int, ...]) -> int:
    """v1 B-tree (node type 1) over chunks [(offsets incl. the trailing 0, nbytes, address)], sorted; -> root address.
    A node is allocated for 2K entries whatever it holds (the library reads whole nodes)."""
    key_size = 8 + 8 * (rank + 1)
    node_size = 24 + 2 * CHUNK_K * (key_size + 8) + key_size

    def key(offs, nbytes):
        return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs)

    level = 0
    nodes = entries                                               # (first key offsets, nbytes of first key, child address)
    while True:
        groups = [nodes[i:i + 2 * CHUNK_K] for i in range(0, len(nodes), 2 * CHUNK_K)] or [[]]
        addrs = [0] * len(groups)
        base = out.put(b"")                                      # nodes of one level are consecutive: siblings are known
        for gi in range(len(groups)):
            addrs[gi] = base + gi * ((node_size + 7) // 8 * 8)
        nxt = []
        for gi, grp in enumerate(groups):
            b = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), addrs[gi - 1] if gi > 0 else UNDEF,
                                      addrs[gi + 1] if gi + 1 < len(groups) else UNDEF)
            for offs, nbytes, child in grp:
                b += key(offs, nbytes) + struct.pack("<Q", child)
            # the closing key: the first key of the right neighbour, or one chunk past the end of the dataset
            if gi + 1 < len(groups):
                b += key(groups[gi + 1][0][0], groups[gi + 1][0][1])
            else:
                b += key(end_key, 0)
            b += b"\0" * (node_size - len(b))
            at = out.put(b)
            assert at == addrs[gi]
            first = grp[0] if grp else (end_key, 0, 0)
            nxt.append((first[0], first[1], at))
        if len(nxt) == 1:
            return nxt[0][2]
        nodes, level = nxt, level + 1


def _write_dataset(out: _Out, arr: np.ndarray, compress: int, earray: bool) -> int:
    arr = np.ascontiguousarray(arr)
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    msgs = [_message(MSG_DATATYPE, _encode_datatype(arr.dtype), 1)]
    fill = _message(MSG_FILL, struct.pack("<BBBBI", 2, 3 if compress else 2, 2, 1, 0), 1)   # incremental / late allocation, fill if set, default value
    if compress and arr.ndim >= 1 and arr.shape[0] > 0 and arr.size > 0:
        rank = arr.ndim
        chunk = (1,) + tuple(arr.shape[1:])
        entries = []
        for i in range(arr.shape[0]):
            raw = zlib.compress(arr[i:i + 1].tobytes(), int(compress))
            entries.append(((i,) + (0,) * rank, len(raw), out.put(raw)))
        root = _chunk_btree(out, entries, rank, (arr.shape[0],) + (0,) * rank)
        msgs.insert(0, _message(MSG_DATASPACE, _dataspace(arr.shape, (None,) + tuple(arr.shape[1:]))))
        msgs.append(fill)
        msgs.append(_message(MSG_FILTERS, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHHI4x", 1, 0, 1, 1, int(compress))))
        lay = struct.pack("<BBBQ", 3, 2, rank + 1, root) + b"".join(struct.pack("<I", c) for c in chunk + (arr.dtype.itemsize,))
        msgs.append(_message(MSG_LAYOUT, lay))
        if earray:                                                 # what PyTables writes on an EArray
            msgs += [_message(MSG_ATTRIBUTE, _attribute("CLASS", "EARRAY")), _message(MSG_ATTRIBUTE, _attribute("EXTDIM", np.int32(0))),
                     _message(MSG_ATTRIBUTE, _attribute("TITLE", "")), _message(MSG_ATTRIBUTE, _attribute("VERSION", "1.1"))]
    else:
        addr = out.put(arr.tobytes()) if arr.size else UNDEF
        msgs.insert(0, _message(MSG_DATASPACE, _dataspace(arr.shape)))
        msgs.append(fill)
        msgs.append(_message(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
    return out.put(_object_header(msgs))


def write(path: str, arrays: Dict[str, np.ndarray], compress: int = 0, earray: bool = True) -> None:
    """Create `path` with one root-level dataset per entry of `arrays`.  compress > 0: chunked along axis 0 (one entry per
    chunk, unlimited) + deflate at that level, the layout of the reference's PyTables EArrays (iotool.py:226-236)."""
    names = sorted(arrays, key=lambda s: s.encode())
    for n in names:
        if not n or "/" in n:
            raise ValueError("h5lite.write: bad dataset name %r" % n)
    if len(names) > 2 * GROUP_LEAF_K * 2 * GROUP_INTERNAL_K:
        raise NotImplementedError("h5lite.write: more than %d datasets" % (4 * GROUP_LEAF_K * GROUP_INTERNAL_K))
    out = _Out()
    out.put(b"\0" * 96)                                            # superblock, filled in last
    headers = [_write_dataset(out, np.asarray(arrays[n]), compress, earray) for n in names]
    # local heap: offset 0 = the empty string, then the names
    heap_data, name_off = bytearray(b"\0" * 8), []
    for n in names:
        name_off.append(len(heap_data))
        heap_data += _pad8(n.encode() + b"\0")
    free_at = len(heap_data)                                      # one free block closes the segment: (next = none, size)
    heap_data += struct.pack("<QQ", HEAP_FREE_NULL, 16)
    heap_data_addr = out.put(bytes(heap_data))
    heap_addr = out.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_at, heap_data_addr))
    # symbol-table nodes (<= 2K entries each, allocated in full), one group B-tree node above them
    snods = []
    per = 2 * GROUP_LEAF_K
    for i in range(0, max(len(names), 1), per):
        part = list(range(i, min(i + per, len(names))))
        b = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
        for j in part:
            b += struct.pack("<QQII16x", name_off[j], headers[j], 0, 0)
        b += b"\0" * (8 + per * 40 - len(b))
        snods.append((out.put(b), name_off[part[-1]] if part else 0))
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
    for addr, last_name in snods:
        bt += struct.pack("<QQ", addr, last_name)
    bt += b"\0" * (24 + (2 * GROUP_INTERNAL_K) * 16 + 8 - len(bt))
    btree_addr = out.put(bt)
    root_msgs = [_message(MSG_SYMBOL_TABLE, struct.pack("<QQ", btree_addr, heap_addr))]
    if earray and compress:                                        # PyTables' root-group attributes
        root_msgs += [_message(MSG_ATTRIBUTE, _attribute("CLASS", "GROUP")), _message(MSG_ATTRIBUTE, _attribute("PYTABLES_FORMAT_VERSION", "2.1")),
                      _message(MSG_ATTRIBUTE, _attribute("TITLE", "")), _message(MSG_ATTRIBUTE, _attribute("VERSION", "1.0"))]
    root_addr = out.put(_object_header(root_msgs))
    out.put(b"")
    eof = out.tell()
    sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == 96
    out.buf[0:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(out.buf))


# ========================================================================================================= reader
class _Obj(object):
    def __init__(self):
        self.msgs: List[Tuple[int, bytes]] = []

    def first(self, mtype: int) -> Optional[bytes]:
        for t, d in self.msgs:
            if t == mtype:
                return d
        return None


class File(object):
    """Read-only view of an HDF5 file: f.keys(), f[name] -> numpy array, f.shape(name), f.attrs(name), "a/b" paths."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("h5lite.File is read-only; use h5lite.write()")
        with open(path, "rb") as f:
            self._b = f.read()
        self._parse_superblock()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self._b = b""

    # ------------------------------------------------------------------ low level
    def _u(self, at: int, n: int) -> int:
        return int.from_bytes(self._b[at:at + n], "little")

    def _addr(self, at: int) -> int:
        v = self._u(at, self._so)
        return UNDEF if v == (1 << (8 * self._so)) - 1 else v + self._base

    def _parse_superblock(self):
        b = self._b
        at = 0
        while b[at:at + 8] != SIGNATURE:                          # a user block may precede it: 512, 1024, ...
            at = 512 if at == 0 else at * 2
            if at + 8 > len(b):
                raise ValueError("h5lite: not an HDF5 file (no signature)")
        ver = b[at + 8]
        self._base = 0
        if ver in (0, 1):
            self._so, self._sl = b[at + 13], b[at + 14]
            p = at + 24 + (4 if ver == 1 else 0)
            self._base = self._u(p, self._so)
            p += 4 * self._so                                      # base, free space, eof, driver
            self._root = ("symtab_entry", p)
        elif ver in (2, 3):
            self._so, self._sl = b[at + 9], b[at + 10]
            p = at + 12
            self._base = self._u(p, self._so)
            self._root = ("header", self._addr(p + 3 * self._so))
        else:
            raise NotImplementedError("h5lite: superblock version %d" % ver)

    def _root_obj(self) -> _Obj:
        kind, v = self._root
        if kind == "header":
            return self._object(v)
        return self._object(self._addr(v + self._so))

    def _object(self, at: int) -> _Obj:
        o = _Obj()
        b = self._b
        if b[at:at + 4] == b"OHDR":                               # version 2
            flags = b[at + 5]
            p = at + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nsz = 1 << (flags & 3)
            size0 = self._u(p, nsz)
            p += nsz
            blocks = [(p, p + size0)]
            track = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, _mflags = b[p], self._u(p + 1, 2), b[p + 3]
                    p += 4 + (2 if track else 0)
                    data = b[p:p + msize]
                    p += msize
                    if mtype == MSG_CONTINUATION:
                        ca, cl = self._addr(p - msize), self._u(p - msize + self._so, self._sl)
                        if b[ca:ca + 4] != b"OCHK":
                            raise ValueError("h5lite: bad object header continuation")
                        blocks.append((ca + 4, ca + cl - 4))
                    elif mtype != 0:
                        o.msgs.append((mtype, data))
            return o
        if b[at] != 1:
            raise NotImplementedError("h5lite: object header version %d" % b[at])
        nmsg, hsize = self._u(at + 2, 2), self._u(at + 8, 4)
        blocks = [(at + 16, at + 16 + hsize)]
        while blocks and len(o.msgs) < nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == MSG_CONTINUATION:
                    ca, cl = self._addr(p - msize), self._u(p - msize + self._so, self._sl)
                    blocks.append((ca, ca + cl))
                elif mtype != 0:
                    o.msgs.append((mtype, data))
        return o

    # ------------------------------------------------------------------ groups
    def _links(self, obj: _Obj) -> Dict[str, int]:
        res: Dict[str, int] = {}
        st = obj.first(MSG_SYMBOL_TABLE)
        if st is not None:
            so = self._so
            btree = int.from_bytes(st[0:so], "little") + self._base
            heap = int.from_bytes(st[so:2 * so], "little") + self._base
            if self._b[heap:heap + 4] != b"HEAP":
                raise ValueError("h5lite: bad local heap")
            heap_data = self._addr(heap + 8 + 2 * self._sl)
            self._walk_group_btree(btree, heap_data, res)
            return res
        for t, d in obj.msgs:
            if t == MSG_LINK:
                flags = d[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = d[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                ln = 1 << (flags & 3)
                nlen = int.from_bytes(d[p:p + ln], "little")
                p += ln
                name = d[p:p + nlen].decode()
                p += nlen
                if ltype != 0:
                    continue                                       # soft / external links are not followed
                res[name] = int.from_bytes(d[p:p + self._so], "little") + self._base
            elif t == MSG_LINK_INFO:
                flags = d[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = int.from_bytes(d[p:p + self._so], "little")
                if fheap != (1 << (8 * self._so)) - 1:
                    raise NotImplementedError("h5lite: dense link storage (fractal heap) -- groups with many links written "
                                              "with libver='latest'")
        return res

    def _walk_group_btree(self, at: int, heap_data: int, res: Dict[str, int]):
        b = self._b
        if b[at:at + 4] == b"SNOD":
            n = self._u(at + 6, 2)
            p = at + 8
            for _ in range(n):
                noff = self._u(p, self._so)
                hdr = self._addr(p + self._so)
                e = b.index(b"\0", heap_data + noff)
                res[b[heap_data + noff:e].decode()] = hdr
                p += 2 * self._so + 8 + 16
            return
        if b[at:at + 4] != b"TREE" or b[at + 4] != 0:
            raise ValueError("h5lite: bad group B-tree node")
        n = self._u(at + 6, 2)
        p = at + 8 + 2 * self._so + self._sl                       # skip key 0
        for _ in range(n):
            self._walk_group_btree(self._addr(p), heap_data, res)
            p += self._so + self._sl

    def _resolve(self, name: str) -> _Obj:
        obj = self._root_obj()
        for part in [s for s in name.split("/") if s]:
            links = self._links(obj)
            if part not in links:
                raise KeyError(name)
            obj = self._object(links[part])
        return obj

    def keys(self, group: str = "/") -> List[str]:
        return sorted(self._links(self._resolve(group)))

    def __contains__(self, name: str) -> bool:
        try:
            self._resolve(name)
            return True
        except KeyError:
            return False

    # ------------------------------------------------------------------ datasets
    def _space(self, obj: _Obj) -> Tuple[int, ...]:
        d = obj.first(MSG_DATASPACE)
        if d is None:
            raise KeyError("not a dataset")
        ver, rank = d[0], d[1]
        p = 8 if ver == 1 else 4
        return tuple(int.from_bytes(d[p + 8 * i:p + 8 * i + self._sl], "little") for i in range(rank))

    def shape(self, name: str) -> Tuple[int, ...]:
        return self._space(self._resolve(name))

    def attrs(self, name: str = "/") -> Dict[str, object]:
        res = {}
        for t, d in self._resolve(name).msgs:
            if t != MSG_ATTRIBUTE:
                continue
            ver = d[0]
            nlen, dlen, slen = struct.unpack_from("<HHH", d, 2)
            p = 8 + (1 if ver == 3 else 0)
            pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
            nm = d[p:p + nlen].split(b"\0")[0].decode()
            p += pad(nlen)
            dt = _decode_datatype(d[p:p + dlen])
            p += pad(dlen)
            sp = d[p:p + slen]
            rank = sp[1]
            off = 8 if sp[0] == 1 else 4
            shp = tuple(int.from_bytes(sp[off + 8 * i:off + 8 * i + self._sl], "little") for i in range(rank))
            p += pad(slen)
            cnt = int(np.prod(shp)) if shp else 1
            val = np.frombuffer(d[p:p + cnt * dt.itemsize], dtype=dt, count=cnt)
            if dt.kind == "S":
                res[nm] = val[0].split(b"\0")[0].decode() if cnt == 1 else [v.decode() for v in val]
            else:
                res[nm] = val.reshape(shp) if shp else val[0]
        return res

    def __getitem__(self, name: str) -> np.ndarray:
        obj = self._resolve(name)
        shape = self._space(obj)
        tb = obj.first(MSG_DATATYPE)
        lay = obj.first(MSG_LAYOUT)
        if tb is None or lay is None:
            raise KeyError("%s is not a dataset" % name)
        dt = _decode_datatype(tb)
        count = int(np.prod(shape)) if shape else 1
        ver, cls = lay[0], lay[1]
        if ver not in (3, 4):
            raise NotImplementedError("h5lite: data layout message version %d" % ver)
        if cls == 0:                                              # compact
            n = int.from_bytes(lay[2:4], "little")
            return np.frombuffer(lay[4:4 + n], dtype=dt, count=count).reshape(shape).copy()
        if cls == 1:                                              # contiguous
            addr = int.from_bytes(lay[2:2 + self._so], "little")
            if addr == (1 << (8 * self._so)) - 1:
                return np.zeros(shape, dtype=dt)                   # never written
            return np.frombuffer(self._b, dtype=dt, count=count, offset=addr + self._base).reshape(shape).copy()
        if cls != 2:
            raise NotImplementedError("h5lite: layout class %d (virtual datasets)" % cls)
        filters = self._filters(obj)
        out = np.zeros(shape, dtype=dt)
        if ver == 3:
            nd = lay[2]
            btree = int.from_bytes(lay[3:3 + self._so], "little")
            p = 3 + self._so
            chunk = tuple(int.from_bytes(lay[p + 4 * i:p + 4 * i + 4], "little") for i in range(nd - 1))
            if btree != (1 << (8 * self._so)) - 1:
                self._walk_chunk_btree(btree + self._base, nd, chunk, dt, filters, out)
            return out
        # version 4 (libver='latest'): only the two trivial chunk indexes
        flags, nd, enc = lay[2], lay[3], lay[4]
        p = 5
        chunk = tuple(int.from_bytes(lay[p + enc * i:p + enc * (i + 1)], "little") for i in range(nd - 1))
        p += enc * nd
        itype = lay[p]
        p += 1
        if itype == 1:                                             # single chunk
            size, mask = int(np.prod(chunk)) * dt.itemsize, 0
            if flags & 0x02:
                size = int.from_bytes(lay[p:p + self._sl], "little")
                mask = int.from_bytes(lay[p + self._sl:p + self._sl + 4], "little")
                p += self._sl + 4
            addr = int.from_bytes(lay[p:p + self._so], "little") + self._base
            self._place(out, (0,) * len(shape), chunk, self._unfilter(self._b[addr:addr + size], filters, mask, dt), dt)
            return out
        if itype == 2:                                             # implicit: chunks back to back, no filters
            addr = int.from_bytes(lay[p:p + self._so], "little") + self._base
            csize = int(np.prod(chunk)) * dt.itemsize
            grid = [(s + c - 1) // c for s, c in zip(shape, chunk)]
            for i, idx in enumerate(np.ndindex(*grid)):
                self._place(out, tuple(a * c for a, c in zip(idx, chunk)), chunk, self._b[addr + i * csize:addr + (i + 1) * csize], dt)
            return out
        raise NotImplementedError("h5lite: chunk index type %d (fixed array / extensible array / v2 B-tree: files written "
                                  "with libver='latest')" % itype)

    def _filters(self, obj: _Obj) -> List[Tuple[int, List[int]]]:
        d = obj.first(MSG_FILTERS)
        if d is None:
            return []
        ver, n = d[0], d[1]
        p = 8 if ver == 1 else 2
        res = []
        for _ in range(n):
            fid = int.from_bytes(d[p:p + 2], "little")
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(d[p:p + 2], "little")
                p += 2
            p += 2                                                 # flags
            ncd = int.from_bytes(d[p:p + 2], "little")
            p += 2
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = [int.from_bytes(d[p + 4 * i:p + 4 * i + 4], "little") for i in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            res.append((fid, cd))
        return res

    def _unfilter(self, raw: bytes, filters, mask: int, dt: np.dtype) -> bytes:
        for i in reversed(range(len(filters))):
            if mask & (1 << i):
                continue
            fid, cd = filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:                                         # shuffle: byte planes -> elements
                size = cd[0] if cd else dt.itemsize
                n = len(raw) // size
                body = np.frombuffer(raw, dtype=np.uint8, count=n * size).reshape(size, n).T.tobytes()
                raw = body + raw[n * size:]
            elif fid == 3:                                         # fletcher32: checksum trails the data
                raw = raw[:-4]
            else:
                raise NotImplementedError("h5lite: filter id %d (only deflate, shuffle, fletcher32)" % fid)
        return raw

    @staticmethod
    def _place(out: np.ndarray, offs, chunk, raw: bytes, dt: np.dtype):
        block = np.frombuffer(raw, dtype=dt, count=int(np.prod(chunk))).reshape(chunk)
        sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, out.shape))
        sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
        out[sl_out] = block[sl_in]

    def _walk_chunk_btree(self, at: int, nd: int, chunk, dt, filters, out: np.ndarray):
        b = self._b
        if b[at:at + 4] != b"TREE" or b[at + 4] != 1:
            raise ValueError("h5lite: bad chunk B-tree node")
        level, n = b[at + 5], self._u(at + 6, 2)
        p = at + 8 + 2 * self._so
        key_size = 8 + 8 * nd
        for _ in range(n):
            nbytes, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = tuple(self._u(p + 8 + 8 * i, 8) for i in range(nd - 1))
            child = self._addr(p + key_size)
            if level > 0:
                self._walk_chunk_btree(child, nd, chunk, dt, filters, out)
            elif all(o < s for o, s in zip(offs, out.shape)):
                self._place(out, offs, chunk, self._unfilter(b[child:child + nbytes], filters, mask, dt), dt)
            p += key_size + self._so


def read(path: str, names=None) -> Dict[str, np.ndarray]:
    with File(path) as f:
        return {n: f[n] for n in (names if names is not None else f.keys())}
