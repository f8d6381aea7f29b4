"""Structural validator for the HDF5 subset libvpm_b200 writes (test infrastructure only).

No HDF5 library exists in this image, so nothing can `H5Fopen` the trajectory files.  Beyond reading the data back
(tests/h5mini.py), this walks EVERY address, size and key the "HDF5 File Format Specification Version 2.0" defines for
the structures in the file and checks the invariants the library relies on when it opens a file:

  superblock     version 0/1, 8-byte offsets/lengths, end-of-file address == file size (the library refuses a file whose
                 stored EOF lies beyond the real one and truncates what lies beyond it), free-space index and driver
                 block undefined, root symbol-table entry: cache type 1 with the B-tree / heap addresses of the root
                 header's symbol-table message
  object headers version 1, message count and header size consistent, every message 8-byte aligned in size, continuation
                 blocks inside the file, required dataset messages present exactly once (dataspace, datatype, layout)
  local heaps    signature / version, data segment inside the file, free list: every block inside the segment, >= 16 bytes,
                 8-byte aligned, no cycles, terminated by 1 (H5HL_FREE_NULL), free blocks do not cover the link names
  group B-trees  node type 0, entries <= 2 * internal K, keys = heap offsets of names in non-decreasing name order,
                 sibling pointers consistent; SNOD: version 1, entries <= 2 * leaf K, sorted by name
  chunk B-trees  node type 1, entries <= 2 * istore K, keys strictly increasing and aligned to the chunk grid, chunk size
                 field == chunk bytes, filter mask 0, internal keys == first key of the child, final key == one chunk past
                 the last, all leaves at the same depth, left/right sibling addresses form the level's linked list
  allocation     every structure (at its FULL allocated size, e.g. B-tree nodes with 2K entries) lies inside [base, EOF) and
                 no two structures overlap

`validate(path)` returns a report (counts, extents) or raises H5FormatError.  It is pinned, like the reader, on a file
written by the real HDF5 library (scipy's MATLAB v7.3 sample): a validator that rejected the library's own output would
be checking my reading of the specification, not the format.
"""
import struct

import numpy as np

import h5mini
from h5mini import UNDEF, H5FormatError


class _Extents:
    def __init__(self, lo, hi):
        self.lo, self.hi, self.items = lo, hi, []

    def add(self, what, start, size):
        if size <= 0:
            raise H5FormatError(f"{what}: non-positive size {size}")
        if start < self.lo or start + size > self.hi:
            raise H5FormatError(f"{what}: [{start}, {start + size}) outside the file [{self.lo}, {self.hi})")
        self.items.append((start, start + size, what))

    def check_disjoint(self):
        it = sorted(self.items)
        for (a0, a1, wa), (b0, b1, wb) in zip(it, it[1:]):
            if b0 < a1:
                raise H5FormatError(f"allocations overlap: {wa} [{a0}, {a1}) and {wb} [{b0}, {b1})")
        return it


def _require(cond, msg):
    if not cond:
        raise H5FormatError(msg)


def _header_extents(f, ext, addr, what):
    """version-1 object header: prefix (16 bytes) + first block, continuation blocks; message-level checks"""
    a = f._abs(addr)
    _require(f.buf[a] == 1 and f.buf[a + 1] == 0, f"{what}: object header version/reserved byte")
    nmsg, refcount, hsize = f._u(a + 2, 2), f._u(a + 4, 4), f._u(a + 8, 4)
    _require(refcount >= 1, f"{what}: object reference count {refcount}")
    _require(addr % 8 == 0, f"{what}: object header address not 8-byte aligned")
    ext.add(f"{what} header", a, 16 + hsize)
    blocks, seen, types = [(a + 16, hsize)], 0, []
    while blocks:
        p, left = blocks.pop(0)
        end = p + left
        while p + 8 <= end and seen < nmsg:
            mtype, msize, mflags = f._u(p, 2), f._u(p + 2, 2), f.buf[p + 4]
            _require(msize % 8 == 0, f"{what}: message 0x{mtype:x} size {msize} not a multiple of 8")
            _require(f.buf[p + 5:p + 8] == b"\0\0\0", f"{what}: reserved bytes of message 0x{mtype:x}")
            _require(p + 8 + msize <= end, f"{what}: message 0x{mtype:x} runs past its block")
            types.append(mtype)
            seen += 1
            if mtype == 0x10:
                caddr, clen = struct.unpack_from("<QQ", f.buf, p + 8)
                ext.add(f"{what} continuation", f._abs(caddr), clen)
                blocks.append((f._abs(caddr), clen))
            p += 8 + msize
        if not blocks:
            # what is left of the last block must be a gap too small for a message or covered by NIL messages
            _require(seen == nmsg, f"{what}: {seen} of {nmsg} messages found")
    return types


def _heap(f, ext, heap_addr, used_offsets):
    h = f._abs(heap_addr)
    _require(f.buf[h:h + 4] == b"HEAP" and f.buf[h + 4] == 0, "local heap signature / version")
    _require(f.buf[h + 5:h + 8] == b"\0\0\0", "local heap reserved bytes")
    dsize, free_head, daddr = struct.unpack_from("<QQQ", f.buf, h + 8)
    ext.add("local heap header", h, 32)
    d = f._abs(daddr)
    ext.add("local heap data", d, dsize)
    _require(dsize % 8 == 0, "local heap data segment size not a multiple of 8")
    # names in use: [offset, end of the NUL-terminated string, padded to 8]
    used = []
    for off in sorted(set(used_offsets)):
        _require(off < dsize, f"heap offset {off} beyond the data segment ({dsize})")
        end = f.buf.index(b"\0", d + off) - d + 1
        _require(end <= dsize, "heap string runs past the data segment")
        used.append((off, (end + 7) // 8 * 8))
    free, seen, nfree = free_head, set(), 0
    while free != 1:       # H5HL_FREE_NULL
        _require(free != UNDEF, "local heap free list ends with the undefined address instead of 1")
        _require(free not in seen, "local heap free list has a cycle")
        seen.add(free)
        _require(free % 8 == 0 and free + 16 <= dsize, f"local heap free block at {free} outside the data segment")
        nxt, size = struct.unpack_from("<QQ", f.buf, d + free)
        _require(size >= 16 and free + size <= dsize, f"local heap free block at {free}: bad size {size}")
        for u0, u1 in used:
            _require(free + size <= u0 or free >= u1, f"local heap free block [{free}, {free + size}) covers the name at {u0}")
        nfree += size
        free = nxt
    return dict(data_size=dsize, free_bytes=nfree, names=len(used))


def _group(f, ext, ohdr, path, report):
    types = _header_extents(f, ext, ohdr, f"group '{path or '/'}'")
    _require(types.count(0x11) == 1, f"group '{path}': symbol table message missing or repeated")
    btree = heap = None
    for mtype, _fl, body, _s in f._messages(ohdr):
        if mtype == 0x11:
            btree, heap = struct.unpack_from("<QQ", f.buf, body)
    name_offsets, children = [], []

    def node(addr, level_expected, left_expected):
        a = f._abs(addr)
        _require(f.buf[a:a + 4] == b"TREE" and f.buf[a + 4] == 0, "group B-tree node signature / type")
        level, used = f.buf[a + 5], f._u(a + 6, 2)
        _require(level_expected is None or level == level_expected, "group B-tree: inconsistent levels")
        _require(used <= 2 * f.internal_k, "group B-tree node over-full")
        ext.add("group B-tree node", a, 24 + (2 * f.internal_k + 1) * 8 + 2 * f.internal_k * 8)
        left, right = struct.unpack_from("<QQ", f.buf, a + 8)
        _require(left == left_expected, "group B-tree: left sibling pointer")
        keys = [f._u(a + 24 + 16 * i, 8) for i in range(used + 1)]
        kids = [f._u(a + 32 + 16 * i, 8) for i in range(used)]
        report["group_nodes"] += 1
        prev = UNDEF
        for i, kid in enumerate(kids):
            if level > 0:
                node(kid, level - 1, prev)
                prev = kid
            else:
                s = f._abs(kid)
                _require(f.buf[s:s + 4] == b"SNOD" and f.buf[s + 4] == 1 and f.buf[s + 5] == 0, "SNOD signature / version")
                n = f._u(s + 6, 2)
                _require(1 <= n <= 2 * f.leaf_k, f"SNOD holds {n} entries (leaf K = {f.leaf_k})")
                ext.add("symbol table node", s, 8 + 2 * f.leaf_k * 40)
                names = []
                for j in range(n):
                    p = s + 8 + 40 * j
                    name_off, child, cache = struct.unpack_from("<QQI", f.buf, p)
                    name_offsets.append(name_off)
                    names.append(f._heap_string(heap, name_off))
                    children.append((names[-1], child, cache, p))
                _require(names == sorted(names) and len(set(names)) == len(names), "SNOD entries not strictly sorted by name")
                # keys bracket the node's names: key[i] < first name <= ... <= last name == key[i+1]
                kl = f._heap_string(heap, keys[i]) if keys[i] != 0 or i > 0 else ""
                kr = f._heap_string(heap, keys[i + 1])
                _require(kl < names[0] or (kl == "" and i == 0), "group B-tree: left key not below the node's first name")
                _require(kr == names[-1], "group B-tree: right key is not the node's last name")
                report["snods"] += 1
        return right

    node(btree, None, UNDEF)
    hrep = _heap(f, ext, heap, name_offsets + [0])
    report["heaps"].append(hrep)
    for name, child, cache, p in children:
        sub = f"{path}/{name}" if path else name
        ctypes = [m[0] for m in f._messages(child)]
        if 0x11 in ctypes:
            _require(cache in (0, 1), f"'{sub}': symbol table entry cache type {cache}")
            if cache == 1:
                cb, ch = struct.unpack_from("<QQ", f.buf, p + 24)
                sb, sh = next(struct.unpack_from("<QQ", f.buf, b) for t, _f2, b, _s2 in f._messages(child) if t == 0x11)
                _require((cb, ch) == (sb, sh), f"'{sub}': cached B-tree / heap addresses differ from the symbol table message")
            _group(f, ext, child, sub, report)
        else:
            _require(cache == 0, f"dataset '{sub}': symbol table entry cache type {cache}")
            _dataset(f, ext, child, sub, report)
    return btree, heap


def _dataset(f, ext, ohdr, path, report):
    types = _header_extents(f, ext, ohdr, f"dataset '{path}'")
    for need, nm in ((0x01, "dataspace"), (0x03, "datatype"), (0x08, "layout")):
        _require(types.count(need) == 1, f"dataset '{path}': {nm} message missing or repeated")
    ds = f.datasets[path]
    esize = ds.dtype.itemsize
    rank = len(ds.shape)
    if ds.maxshape is not None:
        for s, m in zip(ds.shape, ds.maxshape):
            _require(m is None or s <= m, f"dataset '{path}': dimension {s} above its maximum {m}")
    if ds.layout == "contiguous":
        if ds.addr != UNDEF:
            nbytes = int(np.prod(ds.shape)) * esize
            _require(getattr(ds, "nbytes", None) in (None, nbytes), f"dataset '{path}': contiguous size field")
            if nbytes:
                ext.add(f"dataset '{path}' data", f._abs(ds.addr), nbytes)
        report["datasets"][path] = dict(layout="contiguous", shape=ds.shape)
        return
    if ds.layout == "compact":
        report["datasets"][path] = dict(layout="compact", shape=ds.shape)
        return
    _require(len(ds.chunk) == rank and ds.elem == esize, f"dataset '{path}': chunk rank / element size")
    _require(all(c >= 1 for c in ds.chunk), f"dataset '{path}': zero chunk dimension")
    csize = int(np.prod(ds.chunk)) * esize
    _require(csize < 2 ** 32, f"dataset '{path}': chunk of {csize} bytes does not fit the 32-bit chunk size field")
    if any(m is None for m in (ds.maxshape or ())):
        _require(True, "")   # unlimited dimensions require chunked layout: satisfied
    nchunks_expected = int(np.prod([-(-s // c) for s, c in zip(ds.shape, ds.chunk)]))
    ndims = rank + 1
    ksize = 8 + 8 * ndims
    node_bytes = 24 + 2 * f.istore_k * 8 + (2 * f.istore_k + 1) * ksize
    chunks, levels = [], {}

    def node(addr, level_expected):
        a = f._abs(addr)
        _require(f.buf[a:a + 4] == b"TREE" and f.buf[a + 4] == 1, f"dataset '{path}': chunk B-tree node signature / type")
        level, used = f.buf[a + 5], f._u(a + 6, 2)
        _require(level_expected is None or level == level_expected, f"dataset '{path}': chunk B-tree leaves at different depths")
        _require(1 <= used <= 2 * f.istore_k, f"dataset '{path}': chunk B-tree node holds {used} entries (K = {f.istore_k})")
        ext.add(f"dataset '{path}' chunk B-tree node", a, node_bytes)
        left, right = struct.unpack_from("<QQ", f.buf, a + 8)
        keys, kids, p = [], [], a + 24
        for i in range(used + 1):
            nbytes, mask = struct.unpack_from("<II", f.buf, p)
            offs = struct.unpack_from(f"<{ndims}Q", f.buf, p + 8)
            keys.append((nbytes, mask, offs))
            p += ksize
            if i < used:
                kids.append(f._u(p, 8))
                p += 8
        offs_only = [k[2] for k in keys]
        _require(offs_only == sorted(offs_only) and len(set(offs_only)) == len(offs_only),
                 f"dataset '{path}': chunk B-tree keys not strictly increasing")
        for nbytes, mask, offs in keys[:-1]:
            _require(mask == 0, f"dataset '{path}': filter mask set on an unfiltered dataset")
            _require(offs[-1] == 0 and all(o % c == 0 for o, c in zip(offs, ds.chunk)), f"dataset '{path}': key {offs} off the chunk grid")
            _require(nbytes == csize, f"dataset '{path}': chunk size field {nbytes} != {csize}")
        levels.setdefault(level, []).append((addr, left, right))
        first_keys = []
        for i, kid in enumerate(kids):
            if level == 0:
                ext.add(f"dataset '{path}' chunk {keys[i][2][:-1]}", f._abs(kid), csize)
                chunks.append(keys[i][2][:-1])
            else:
                cf, cl = node(kid, level - 1)
                _require(cf == keys[i][2], f"dataset '{path}': internal key differs from the child's first key")
                _require(cl == keys[i + 1][2], f"dataset '{path}': internal key differs from the child's last key")
        return keys[0][2], keys[-1][2]

    if ds.addr != UNDEF:
        _first, last = node(ds.addr, None)
        for level, nodes in levels.items():      # the nodes of one level, in key order, are a doubly linked list
            for i, (addr, left, right) in enumerate(nodes):
                _require(left == (nodes[i - 1][0] if i else UNDEF), f"dataset '{path}': level {level} left sibling pointer")
                _require(right == (nodes[i + 1][0] if i + 1 < len(nodes) else UNDEF), f"dataset '{path}': level {level} right sibling pointer")
        # the final key is one chunk past the last one along the slowest dimension
        lastchunk = chunks[-1]
        _require(tuple(last[:-1]) > tuple(lastchunk), f"dataset '{path}': final key {last} not beyond the last chunk {lastchunk}")
    _require(len(set(chunks)) == len(chunks), f"dataset '{path}': chunk indexed twice")
    for c in chunks:
        _require(all(o < s for o, s in zip(c, ds.shape)), f"dataset '{path}': chunk {c} outside the dataspace {ds.shape}")
    report["datasets"][path] = dict(layout="chunked", shape=ds.shape, chunk=ds.chunk, chunks=len(chunks), chunks_possible=nchunks_expected,
                                    btree_depth=(max(levels) + 1 if levels else 0), btree_nodes=sum(len(v) for v in levels.values()))


def validate(path, require_exact_eof=True):
    f = h5mini.File(path)
    import os
    fsize = os.path.getsize(path)
    b = f.base
    _require(f.buf[b + 9] == 0 and f.buf[b + 10] == 0 and f.buf[b + 12] == 0, "superblock: free-space / root-group / shared-header versions")
    _require(f.buf[b + 11] == 0 and f.buf[b + 15] == 0, "superblock: reserved bytes")
    _require(f.leaf_k >= 1 and f.internal_k >= 1, "superblock: zero B-tree K")
    _require(f.base_addr in (0, UNDEF) or f.base_addr == b, "superblock: base address")
    _require(f.free_addr == UNDEF, "superblock: global free-space index must be undefined")
    _require(f.driver_addr == UNDEF, "superblock: driver information block must be undefined")
    # the library stores base address + relative end-of-allocation here, i.e. an ABSOLUTE file offset (seen in the
    # libhdf5-written sample: base 512, EOF field 4168 = the file's size)
    eof_abs = f.eof_addr
    if require_exact_eof:
        _require(eof_abs == fsize, f"superblock: end-of-file address {eof_abs} != file size {fsize}")
    else:
        _require(eof_abs <= fsize, f"superblock: end-of-file address {eof_abs} beyond the file ({fsize})")
    ext = _Extents(b, eof_abs)
    ext.add("superblock", b, f.superblock_size)
    report = dict(file_size=fsize, eof=eof_abs, base=b, group_nodes=0, snods=0, heaps=[], datasets={})
    re_ = f.root_entry
    _require(re_["name_off"] == 0, "root symbol table entry: link name offset must be 0")
    btree, heap = _group(f, ext, re_["ohdr"], "", report)
    _require(re_["cache"] in (0, 1), f"root symbol table entry: cache type {re_['cache']}")
    if re_["cache"] == 1:
        _require((re_["btree"], re_["heap"]) == (btree, heap), "root symbol table entry: cached B-tree / heap addresses")
    allocs = ext.check_disjoint()
    report["allocations"] = len(allocs)
    report["allocated_bytes"] = sum(a1 - a0 for a0, a1, _w in allocs)
    return report
