"""Minimal pure-Python HDF5 reader / writer.

The image has no libhdf5, h5py or PyTables, yet the reference's file formats are part of its API
surface: the AMISR fitted input file (`Interpolate.read_datafile`, reference interpolate.py:582-667)
and the coefficient output file (`saveh5`, interpolate.py:671-708; read back by `Estimate.loadh5`,
estimate.py:53-70).  This module implements the subset of the HDF5 file format (spec version 1.x:
superblock v0/v1, v1 object headers with continuation blocks, symbol-table groups = v1 B-tree + local
heap + SNOD, contiguous and chunked (v1 B-tree, deflate / shuffle / fletcher32) layouts, fixed-point,
IEEE float and fixed-length string datatypes, v1 attributes) that those files use.

Writer: contiguous datasets only (what PyTables `create_array` produces), plus the node attributes
PyTables sets (CLASS / VERSION / TITLE / FLAVOR) so that the reference's own `Estimate` can open the
file.  NOTE: written from the published format specification; it could not be cross-checked against
libhdf5 in this image.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


# ================================================================================================
# reader
# ================================================================================================
class File(object):
    """h5[path] -> numpy array (datasets) ; h5.attrs(path) -> dict ; h5.keys(path) -> member names."""

    def __init__(self, filename):
        with open(filename, "rb") as f:
            self.buf = f.read()
        self._parse_superblock()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # ---- low level -----------------------------------------------------------------------------
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _parse_superblock(self):
        b = self.buf
        base = 0
        while b[base:base + 8] != SIG:
            base = 512 if base == 0 else base * 2
            if base >= len(b):
                raise ValueError("not an HDF5 file")
        ver = b[base + 8]
        if ver not in (0, 1):
            raise NotImplementedError("superblock version %d (only 0/1: libver='earliest' files)" % ver)
        self.so, self.sl = b[base + 13], b[base + 14]
        if (self.so, self.sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets/lengths")
        off = base + 24 + (4 if ver == 1 else 0)
        self.base = self._u(off, 8)
        root = off + 32                      # root group symbol table entry
        self.root_header = self._u(root + 8, 8) + self.base

    def _messages(self, addr):
        """All (type, flags, data offset, size) of a version-1 object header incl. continuation blocks."""
        b = self.buf
        if b[addr] != 1:
            raise NotImplementedError("object header version %d (only v1)" % b[addr])
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                data = p + 8
                if mtype == 0x0010:          # continuation
                    blocks.append((self._u(data, 8) + self.base, self._u(data + 8, 8)))
                out.append((mtype, flags, data, msize))
                p = data + msize
        return out

    def _group_members(self, header):
        for mtype, _, data, _ in self._messages(header):
            if mtype == 0x0011:
                btree, heap = self._u(data, 8) + self.base, self._u(data + 8, 8) + self.base
                if self.buf[heap:heap + 4] != b"HEAP":
                    raise ValueError("bad local heap")
                hdata = self._u(heap + 24, 8) + self.base
                out = {}
                self._walk_group_btree(btree, hdata, out)
                return out
        raise KeyError("not a group (no symbol table message)")

    def _walk_group_btree(self, node, hdata, out):
        b = self.buf
        if b[node:node + 4] == b"SNOD":
            n = self._u(node + 6, 2)
            for i in range(n):
                e = node + 8 + 40 * i
                noff = self._u(e, 8)
                end = b.index(b"\x00", hdata + noff)
                out[b[hdata + noff:end].decode("utf-8")] = self._u(e + 8, 8) + self.base
            return
        if b[node:node + 4] != b"TREE":
            raise ValueError("bad group B-tree node")
        used = self._u(node + 6, 2)
        p = node + 24
        for i in range(used):
            child = self._u(p + 8 + 16 * i, 8) + self.base
            self._walk_group_btree(child, hdata, out)

    def _resolve(self, path):
        addr = self.root_header
        for part in [p for p in path.split("/") if p]:
            addr = self._group_members(addr)[part]
        return addr

    def keys(self, path="/"):
        return sorted(self._group_members(self._resolve(path)))

    # ---- datatypes / dataspaces -------------------------------------------------------------------
    def _dtype(self, p):
        b = self.buf
        cls, bits0, size = b[p] & 0x0F, b[p + 1], self._u(p + 4, 4)
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if (bits0 & 0x08) else "u", size))
        if cls == 1:
            return np.dtype("%sf%d" % (order, size))
        if cls == 3:
            return np.dtype("S%d" % size)
        raise NotImplementedError("datatype class %d" % cls)

    def _shape(self, p):
        ver, rank = self.buf[p], self.buf[p + 1]
        start = p + (8 if ver == 1 else 4)
        return tuple(self._u(start + 8 * i, 8) for i in range(rank))

    def attrs(self, path):
        out = {}
        for mtype, _, data, _ in self._messages(self._resolve(path)):
            if mtype != 0x000C:
                continue
            ver = self.buf[data]
            nsz, tsz, ssz = self._u(data + 2, 2), self._u(data + 4, 2), self._u(data + 6, 2)
            pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
            p = data + 8
            name = self.buf[p:p + nsz].split(b"\x00")[0].decode("utf-8")
            p += pad(nsz)
            dt = self._dtype(p)
            p += pad(tsz)
            shape = self._shape(p)
            p += pad(ssz)
            n = int(np.prod(shape)) if shape else 1
            val = np.frombuffer(self.buf, dtype=dt, count=n, offset=p).reshape(shape)
            out[name] = val[()] if not shape else val.copy()
        return out

    # ---- datasets -----------------------------------------------------------------------------------
    def __getitem__(self, path):
        header = self._resolve(path)
        dt = shape = layout = None
        filters = []
        for mtype, _, data, size in self._messages(header):
            if mtype == 0x0001:
                shape = self._shape(data)
            elif mtype == 0x0003:
                dt = self._dtype(data)
            elif mtype == 0x0008:
                layout = data
            elif mtype == 0x000B:
                filters = self._filters(data)
        if dt is None or shape is None or layout is None:
            raise KeyError("%s is not a dataset" % path)
        b = self.buf
        ver, cls = b[layout], b[layout + 1]
        if ver != 3:
            raise NotImplementedError("data layout message version %d" % ver)
        n = int(np.prod(shape)) if shape else 1
        if cls == 1:                         # contiguous
            addr = self._u(layout + 2, 8)
            if addr == UNDEF:
                arr = np.zeros(shape, dtype=dt)
            else:
                arr = np.frombuffer(b, dtype=dt, count=n, offset=addr + self.base).reshape(shape).copy()
        elif cls == 0:                       # compact
            sz = self._u(layout + 2, 2)
            arr = np.frombuffer(b[layout + 4:layout + 4 + sz], dtype=dt, count=n).reshape(shape).copy()
        elif cls == 2:                       # chunked
            rank1 = b[layout + 2]
            btree = self._u(layout + 3, 8)
            cdims = tuple(self._u(layout + 11 + 4 * i, 4) for i in range(rank1))[:-1]
            arr = np.zeros(shape, dtype=dt)
            if btree != UNDEF:
                self._read_chunks(btree + self.base, arr, cdims, dt, filters)
        else:
            raise NotImplementedError("layout class %d" % cls)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        if dt.kind == "S" and arr.shape == ():
            return bytes(arr[()])
        return arr

    def _filters(self, p):
        ver, nf = self.buf[p], self.buf[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(nf):
            fid = self._u(q, 2)
            if ver == 1 or fid >= 256:
                nlen = self._u(q + 2, 2)
                flags, ncd = self._u(q + 4, 2), self._u(q + 6, 2)
                q += 8 + ((nlen + 7) // 8 * 8 if ver == 1 else nlen)
            else:
                flags, ncd = self._u(q + 2, 2), self._u(q + 4, 2)
                q += 6
            cd = [self._u(q + 4 * i, 4) for i in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _read_chunks(self, node, arr, cdims, dt, filters):
        b = self.buf
        if b[node:node + 4] != b"TREE":
            raise ValueError("bad chunk B-tree node")
        level, used = b[node + 5], self._u(node + 6, 2)
        rank = len(cdims)
        ksz = 8 + 8 * (rank + 1)
        p = node + 24
        for i in range(used):
            key = p + i * (ksz + 8)
            csize, mask = self._u(key, 4), self._u(key + 4, 4)
            offs = tuple(self._u(key + 8 + 8 * d, 8) for d in range(rank))
            child = self._u(key + ksz, 8) + self.base
            if level > 0:
                self._read_chunks(child, arr, cdims, dt, filters)
                continue
            raw = b[child:child + csize]
            for j, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << j):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dt.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    nel = a.size // es
                    raw = a[:nel * es].reshape(es, nel).T.tobytes() + a[nel * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise NotImplementedError("HDF5 filter %d" % fid)
            chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, arr.shape))
            arr[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]


# ================================================================================================
# writer
# ================================================================================================
def _pad8(bts):
    return bts + b"\x00" * (-len(bts) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        if dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + props
        props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + props
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))     # null-padded ASCII
    raise TypeError("unsupported dtype %r" % dt)


def _space_msg(shape):
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attr_msg(name, value):
    if isinstance(value, (bytes, str)):
        raw = value.encode("utf-8") if isinstance(value, str) else value
        arr = np.array(raw if raw else b"\x00", dtype="S%d" % max(len(raw), 1))
    else:
        arr = np.asarray(value)
    nm = name.encode("utf-8") + b"\x00"
    dtm, spm = _dtype_msg(arr.dtype), _space_msg(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(spm)) + _pad8(nm) + _pad8(dtm) + _pad8(spm) + arr.tobytes()
    return _msg(0x000C, body)


class Writer(object):
    """Builds the whole file in memory, writes it on close.  Groups use symbol tables (one SNOD per
    group: up to 2*LEAF_K members), datasets are contiguous."""

    LEAF_K, INTERNAL_K = 32, 16

    def __init__(self, filename, pytables_attrs=True):
        self.filename = filename
        self.pt = pytables_attrs
        self.tree = {"kind": "group", "title": "", "members": {}}

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        if exc_type is None:
            self.close()
        return False

    def _parent(self, path, create=True):
        parts = [p for p in path.split("/") if p]
        node = self.tree
        for p in parts[:-1]:
            if p not in node["members"]:
                if not create:
                    raise KeyError(path)
                node["members"][p] = {"kind": "group", "title": "", "members": {}}
            node = node["members"][p]
        return node, parts[-1]

    def group(self, path, title=""):
        parent, name = self._parent(path)
        parent["members"].setdefault(name, {"kind": "group", "title": title, "members": {}})["title"] = title

    def array(self, path, data):
        """numeric ndarray -> contiguous dataset (PyTables create_array of an ndarray: FLAVOR 'numpy')."""
        parent, name = self._parent(path)
        a = np.ascontiguousarray(data)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        parent["members"][name] = {"kind": "data", "arr": a, "flavor": "numpy"}

    def string(self, path, value):
        """bytes scalar -> rank-0 fixed-length string (PyTables create_array of bytes: FLAVOR 'python')."""
        parent, name = self._parent(path)
        raw = bytes(value)
        parent["members"][name] = {"kind": "data", "arr": np.array(raw if raw else b"\x00", dtype="S%d" % max(len(raw), 1)),
                                   "flavor": "python"}

    def strings(self, path, values):
        """list of str -> 1-D fixed-length string array (FLAVOR 'python')."""
        parent, name = self._parent(path)
        vals = [v.encode("utf-8") if isinstance(v, str) else bytes(v) for v in values]
        n = max([len(v) for v in vals] + [1])
        parent["members"][name] = {"kind": "data", "arr": np.array(vals, dtype="S%d" % n).reshape(len(vals)),
                                   "flavor": "python"}

    # ---- serialisation -----------------------------------------------------------------------------
    def close(self):
        self.out = bytearray(96)             # superblock placeholder
        root_hdr, root_btree, root_heap = self._emit_group(self.tree, root=True)
        eof = len(self.out)
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
        assert len(sb) == 96
        self.out[0:96] = sb
        with open(self.filename, "wb") as f:
            f.write(bytes(self.out))

    def _alloc(self, bts):
        while len(self.out) % 8:
            self.out.append(0)
        addr = len(self.out)
        self.out += bts
        return addr

    def _header(self, msgs):
        body = b"".join(msgs)
        return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body

    def _emit_data(self, node):
        a = node["arr"]
        addr = self._alloc(a.tobytes()) if a.size else UNDEF
        msgs = [_msg(0x0001, _space_msg(a.shape)), _msg(0x0003, _dtype_msg(a.dtype), flags=1),
                _msg(0x0005, struct.pack("<BBBB", 2, 1, 0, 0)),
                _msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, a.nbytes))]
        if self.pt:
            msgs += [_attr_msg("CLASS", b"ARRAY"), _attr_msg("VERSION", b"2.4"), _attr_msg("TITLE", b""),
                     _attr_msg("FLAVOR", node["flavor"].encode())]
        return self._alloc(self._header(msgs))

    def _emit_group(self, node, root=False):
        names = sorted(node["members"], key=lambda s: s.encode("utf-8"))
        if len(names) > 2 * self.LEAF_K:
            raise NotImplementedError("more than %d members in one group" % (2 * self.LEAF_K))
        entries = []
        for nm in names:
            m = node["members"][nm]
            if m["kind"] == "group":
                hdr, bt, hp = self._emit_group(m)
                entries.append((nm, hdr, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((nm, self._emit_data(m), 0, b"\x00" * 16))
        # local heap: "" at offset 0, then the member names
        heap = bytearray(8)
        offs = []
        for nm, *_ in entries:
            offs.append(len(heap))
            heap += _pad8(nm.encode("utf-8") + b"\x00")
        heap_data = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))
        # symbol table node
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(entries)))
        for (nm, hdr, cache, scratch), off in zip(entries, offs):
            snod += struct.pack("<QQII", off, hdr, cache, 0) + scratch
        snod += b"\x00" * (8 + 40 * 2 * self.LEAF_K - len(snod))
        snod_addr = self._alloc(bytes(snod))
        # B-tree: one level-0 node pointing at the SNOD (keys: "" and the largest name)
        bt = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, UNDEF, UNDEF))
        bt += struct.pack("<Q", 0)
        if entries:
            bt += struct.pack("<QQ", snod_addr, offs[-1])
        bt += b"\x00" * (24 + (2 * self.INTERNAL_K + 1) * 8 + 2 * self.INTERNAL_K * 8 - len(bt))
        bt_addr = self._alloc(bytes(bt))
        msgs = [_msg(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
        if self.pt:
            msgs += [_attr_msg("CLASS", b"GROUP"), _attr_msg("VERSION", b"1.0"), _attr_msg("TITLE", node["title"].encode())]
            if root:
                msgs.append(_attr_msg("PYTABLES_FORMAT_VERSION", b"2.1"))
        return self._alloc(self._header(msgs)), bt_addr, heap_addr
