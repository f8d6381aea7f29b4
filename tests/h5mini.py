"""Minimal, independent HDF5 reader (test infrastructure only).

h5py / HDF5.jl are not in this image, so the files written by libvpm_b200's trajectory writer
(`vpm_h5_*`, csrc/h5min.cpp) are checked with this from-scratch reader of the subset of the HDF5 file
format they use: superblock version 0, old-style groups (symbol-table message, version-1 B-tree of
type 0, local heap, SNOD nodes), version-1 object headers (with continuation blocks), dataspace
version 1, fixed/floating-point datatypes version 1, data layout version 3 (compact, contiguous,
chunked with a version-1 B-tree of type 1), no filters.  The reader itself is pinned on a file written
by the real HDF5 library that ships inside scipy's test data (tests/test_h5_cpu.py).

Written from the published "HDF5 File Format Specification Version 2.0" structures; shares no code
with the writer.
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class Dataset:
    def __init__(self, name):
        self.name = name
        self.shape = None
        self.maxshape = None
        self.dtype = None
        self.layout = None       # "compact" | "contiguous" | "chunked"
        self.chunk = None        # chunk dims (without the element-size entry)
        self.addr = None         # contiguous: data address; chunked: B-tree address
        self.raw = None          # compact data
        self.filters = False
        self.messages = []       # (type, size) of every header message, for inspection
        self.payloads = {}       # message type -> raw payload bytes
        self.chunks = []         # (offsets, address, nbytes) of every chunk found in the B-tree


class File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.base = self._find_superblock()
        self.datasets = {}
        self.group_info = {}
        self._parse_superblock()

    # ---------------------------------------------------------------- primitives
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _find_superblock(self):
        off = 0
        while off < len(self.buf):
            if self.buf[off:off + 8] == SIG:
                return off
            off = 512 if off == 0 else off * 2
        raise H5FormatError("no HDF5 signature")

    def _parse_superblock(self):
        b = self.base
        ver = self.buf[b + 8]
        if ver not in (0, 1):
            raise H5FormatError(f"superblock version {ver} not supported")
        self.sb_version = ver
        self.size_offsets, self.size_lengths = self.buf[b + 13], self.buf[b + 14]
        if (self.size_offsets, self.size_lengths) != (8, 8):
            raise H5FormatError("only 8-byte offsets/lengths")
        self.leaf_k, self.internal_k = self._u(b + 16, 2), self._u(b + 18, 2)
        self.flags = self._u(b + 20, 4)
        p = b + 24
        self.istore_k = 32
        if ver == 1:
            self.istore_k = self._u(p, 2)
            p += 4
        self.base_addr, self.free_addr, self.eof_addr, self.driver_addr = struct.unpack_from("<4Q", self.buf, p)
        p += 32
        # the addresses in the file are relative to the base address
        if self.eof_addr != UNDEF and self.eof_addr > len(self.buf):
            raise H5FormatError(f"end-of-file address {self.eof_addr} beyond the file ({len(self.buf)} bytes)")
        name_off, ohdr, cache, _res = struct.unpack_from("<QQII", self.buf, p)
        self.root_entry = dict(name_off=name_off, ohdr=ohdr, cache=cache)
        if cache == 1:
            self.root_entry["btree"], self.root_entry["heap"] = struct.unpack_from("<QQ", self.buf, p + 24)
        self.superblock_size = p + 40 - b
        self._read_group(ohdr, "")

    def _abs(self, addr):
        # file addresses are relative to the superblock's base address (the user-block size)
        return addr + (self.base_addr if self.base_addr != UNDEF else self.base)

    # ---------------------------------------------------------------- object headers
    def _messages(self, addr):
        """yield (type, flags, payload offset, size) of a version-1 object header, following continuations"""
        a = self._abs(addr)
        if self.buf[a] != 1:
            raise H5FormatError(f"object header version {self.buf[a]} at {addr}")
        nmsg = self._u(a + 2, 2)
        hsize = self._u(a + 8, 4)
        blocks = [(a + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), self.buf[p + 4]
                body = p + 8
                if body + msize > end:
                    raise H5FormatError("header message runs past its block")
                out.append((mtype, mflags, body, msize))
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", self.buf, body)
                    blocks.append((self._abs(caddr), clen))
                p = body + msize
        if len(out) != nmsg:
            raise H5FormatError(f"object header at {addr}: found {len(out)} of {nmsg} messages")
        return out

    # ---------------------------------------------------------------- groups
    def _heap_string(self, heap_addr, off):
        h = self._abs(heap_addr)
        if self.buf[h:h + 4] != b"HEAP":
            raise H5FormatError("local heap signature missing")
        dsize, free_head, daddr = struct.unpack_from("<QQQ", self.buf, h + 8)
        self.group_info.setdefault("heaps", {})[heap_addr] = dict(data_size=dsize, free_head=free_head, data_addr=daddr)
        d = self._abs(daddr) + off
        end = self.buf.index(b"\0", d)
        if end - self._abs(daddr) >= dsize:
            raise H5FormatError("heap string runs past the data segment")
        return self.buf[d:end].decode()

    def _group_nodes(self, btree_addr, heap_addr):
        """symbol-table entries below a type-0 B-tree node, in key order"""
        a = self._abs(btree_addr)
        if self.buf[a:a + 4] != b"TREE" or self.buf[a + 4] != 0:
            raise H5FormatError("group B-tree node expected")
        level, used = self.buf[a + 5], self._u(a + 6, 2)
        entries = []
        p = a + 24
        for i in range(used):
            child = self._u(p + 8, 8)
            if level > 0:
                entries += self._group_nodes(child, heap_addr)
            else:
                entries += self._snod(child, heap_addr)
            p += 16
        return entries

    def _snod(self, addr, heap_addr):
        a = self._abs(addr)
        if self.buf[a:a + 4] != b"SNOD":
            raise H5FormatError("SNOD signature missing")
        n = self._u(a + 6, 2)
        out = []
        for i in range(n):
            p = a + 8 + 40 * i
            name_off, ohdr, cache = struct.unpack_from("<QQI", self.buf, p)
            out.append((self._heap_string(heap_addr, name_off), ohdr, cache, p))
        names = [e[0] for e in out]
        if names != sorted(names):
            raise H5FormatError("symbol table node entries are not sorted by name")
        return out

    def _read_group(self, ohdr, prefix):
        btree = heap = None
        for mtype, _f, body, _s in self._messages(ohdr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self.buf, body)
        if btree is None:
            raise H5FormatError("group without a symbol table message")
        for name, child, cache, p in self._group_nodes(btree, heap):
            path = f"{prefix}/{name}" if prefix else name
            types = [m[0] for m in self._messages(child)]
            if 0x11 in types:
                self._read_group(child, path)
            elif 0x08 in types:
                self.datasets[path] = self._read_dataset(child, path)

    # ---------------------------------------------------------------- datasets
    def _read_dataset(self, ohdr, name):
        ds = Dataset(name)
        for mtype, mflags, body, msize in self._messages(ohdr):
            ds.messages.append((mtype, msize))
            ds.payloads[mtype] = bytes(self.buf[body:body + msize])
            if mtype == 0x01:
                ver, rank, flags = self.buf[body], self.buf[body + 1], self.buf[body + 2]
                if ver != 1:
                    raise H5FormatError(f"dataspace version {ver}")
                p = body + 8
                ds.shape = struct.unpack_from(f"<{rank}Q", self.buf, p)
                if flags & 1:
                    ds.maxshape = tuple(None if m == UNDEF else m for m in struct.unpack_from(f"<{rank}Q", self.buf, p + 8 * rank))
                else:
                    ds.maxshape = ds.shape
            elif mtype == 0x03:
                ds.dtype = self._datatype(body)
            elif mtype == 0x0B:
                ds.filters = True
            elif mtype == 0x08:
                ver, cls = self.buf[body], self.buf[body + 1]
                if ver in (1, 2):  # pre-1.6.3 files (e.g. the library-written sample the reader is pinned on)
                    nd, cls = self.buf[body + 1], self.buf[body + 2]
                    p = body + 8
                    if cls != 0:
                        ds.addr = self._u(p, 8)
                        p += 8
                    dims = struct.unpack_from(f"<{nd}I", self.buf, p)
                    p += 4 * nd
                    if cls == 2:
                        ds.layout, ds.chunk, ds.elem = "chunked", dims[:-1], dims[-1]
                    elif cls == 1:
                        ds.layout, ds.nbytes = "contiguous", None
                    else:
                        n = self._u(p, 4)
                        ds.layout, ds.raw = "compact", self.buf[p + 4:p + 4 + n]
                    continue
                if ver != 3:
                    raise H5FormatError(f"layout version {ver}")
                if cls == 0:
                    n = self._u(body + 2, 2)
                    ds.layout, ds.raw = "compact", self.buf[body + 4:body + 4 + n]
                elif cls == 1:
                    ds.layout = "contiguous"
                    ds.addr, ds.nbytes = struct.unpack_from("<QQ", self.buf, body + 2)
                elif cls == 2:
                    nd = self.buf[body + 2]
                    ds.layout = "chunked"
                    ds.addr = self._u(body + 3, 8)
                    dims = struct.unpack_from(f"<{nd}I", self.buf, body + 11)
                    ds.chunk, ds.elem = dims[:-1], dims[-1]
                else:
                    raise H5FormatError(f"layout class {cls}")
        return ds

    def _datatype(self, body):
        cv = self.buf[body]
        cls, ver = cv & 0xF, cv >> 4
        bits = self._u(body + 1, 3)
        size = self._u(body + 4, 4)
        if ver not in (1, 2, 3):
            raise H5FormatError(f"datatype version {ver}")
        order = ">" if bits & 1 else "<"
        if cls == 0:
            signed = bool(bits & 8)
            prec = self._u(body + 10, 2)
            if prec != 8 * size:
                raise H5FormatError("padded integers not supported")
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:
            off, prec = self._u(body + 8, 2), self._u(body + 10, 2)
            eloc, esize, mloc, msize = self.buf[body + 12:body + 16]
            bias = self._u(body + 16, 4)
            sign = (bits >> 8) & 0xFF
            ieee = {8: (0, 64, 52, 11, 0, 52, 1023, 63), 4: (0, 32, 23, 8, 0, 23, 127, 31)}.get(size)
            if (off, prec, eloc, esize, mloc, msize, bias, sign) != ieee:
                raise H5FormatError("not an IEEE float")
            if (bits >> 4) & 3 != 2:
                raise H5FormatError("mantissa normalisation is not 'implied msb'")
            return np.dtype(f"{order}f{size}")
        return np.dtype(f"V{size}")

    def _chunk_tree(self, addr, ndims, out, level_expected=None):
        a = self._abs(addr)
        if self.buf[a:a + 4] != b"TREE" or self.buf[a + 4] != 1:
            raise H5FormatError("chunk B-tree node expected")
        level, used = self.buf[a + 5], self._u(a + 6, 2)
        if level_expected is not None and level != level_expected:
            raise H5FormatError("B-tree levels inconsistent")
        if used > 2 * self.istore_k:
            raise H5FormatError("B-tree node over-full")
        left, right = struct.unpack_from("<QQ", self.buf, a + 8)
        ksize = 8 + 8 * ndims
        node_bytes = 24 + 2 * self.istore_k * 8 + (2 * self.istore_k + 1) * ksize
        if a + node_bytes > len(self.buf):
            raise H5FormatError("B-tree node (full allocated size) runs past the end of the file")
        p = a + 24
        keys, kids = [], []
        for i in range(used + 1):
            nbytes, mask = struct.unpack_from("<II", self.buf, p)
            offs = struct.unpack_from(f"<{ndims}Q", self.buf, p + 8)
            keys.append((nbytes, mask, offs))
            p += ksize
            if i < used:
                kids.append(self._u(p, 8))
                p += 8
        if [k[2] for k in keys] != sorted(k[2] for k in keys) or len(set(k[2] for k in keys)) != len(keys):
            raise H5FormatError("B-tree keys not strictly increasing")
        info = dict(addr=addr, level=level, used=used, left=left, right=right, first=keys[0], last=keys[-1])
        self.group_info.setdefault("chunk_nodes", []).append(info)
        for i, kid in enumerate(kids):
            if level == 0:
                if keys[i][1] != 0:
                    raise H5FormatError("filtered chunk")
                out.append((keys[i][2], kid, keys[i][0]))
            else:
                sub = []
                first = self._chunk_tree(kid, ndims, sub, level - 1)
                if first["first"][2] != keys[i][2]:
                    raise H5FormatError("internal key differs from the child's first key")
                if first["last"][2] != keys[i + 1][2]:
                    raise H5FormatError("internal key differs from the child's last key")
                out += sub
        return info

    def read(self, name):
        ds = self.datasets[name]
        if ds.filters:
            raise H5FormatError("filtered datasets not supported")
        count = int(np.prod(ds.shape)) if ds.shape else 1
        if ds.layout == "compact":
            return np.frombuffer(ds.raw, ds.dtype, count).reshape(ds.shape)
        if ds.layout == "contiguous":
            if ds.addr == UNDEF:
                return np.zeros(ds.shape, ds.dtype)
            return np.frombuffer(self.buf, ds.dtype, count, self._abs(ds.addr)).reshape(ds.shape)
        out = np.zeros(ds.shape, ds.dtype)
        seen = np.zeros([-(-s // c) for s, c in zip(ds.shape, ds.chunk)], bool)
        ds.chunks = []
        if ds.addr != UNDEF:
            self._chunk_tree(ds.addr, len(ds.shape) + 1, ds.chunks)
        csize = int(np.prod(ds.chunk)) * ds.dtype.itemsize
        for offs, addr, nbytes in ds.chunks:
            if offs[-1] != 0 or nbytes != csize or any(o % c for o, c in zip(offs, ds.chunk)):
                raise H5FormatError(f"bad chunk key {offs} / {nbytes}")
            idx = tuple(o // c for o, c in zip(offs, ds.chunk))
            if seen[idx]:
                raise H5FormatError("chunk listed twice")
            seen[idx] = True
            a = self._abs(addr)
            if a + nbytes > len(self.buf):
                raise H5FormatError("chunk beyond the end of the file")
            block = np.frombuffer(self.buf, ds.dtype, csize // ds.dtype.itemsize, a).reshape(ds.chunk)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, ds.chunk, ds.shape))
            out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out
