"""Trajectory files: libvpm_b200's own HDF5 writer (vpm_h5_*, csrc/h5min.cpp) reproduces the
layout run! writes (src/methods/splitting.jl:32-34, src/methods/geometric_integrator.jl:21-25).  No HDF5
library exists in this image, so the files are read back with tests/h5mini.py, an independent reader that is
first pinned on a file written by the real HDF5 library (scipy ships MATLAB's v7.3 = HDF5 sample next to the
same variable in the classic format, which scipy itself can read)."""
import os

import numpy as np
import pytest

import h5mini

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vpm():
    import __graft_entry__ as g
    g.build()
    import vpm_b200
    return vpm_b200


def test_reader_is_pinned_on_a_file_written_by_libhdf5():
    import scipy.io
    d = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data")
    h5, classic = os.path.join(d, "testhdf5_7.4_GLNX86.mat"), os.path.join(d, "testdouble_7.4_GLNX86.mat")
    if not (os.path.exists(h5) and os.path.exists(classic)):
        pytest.skip("scipy's MATLAB sample files are not installed")
    f = h5mini.File(h5)
    assert f.base == 512 and f.sb_version == 0 and (f.leaf_k, f.internal_k) == (4, 16)
    assert list(f.datasets) == ["testdouble"]
    ds = f.datasets["testdouble"]
    assert ds.dtype == np.dtype("<f8") and ds.layout == "contiguous"
    want = scipy.io.loadmat(classic)["testdouble"]          # (1, 9), MATLAB column-major
    got = f.read("testdouble")                              # HDF5 row-major view of the same array: (9, 1)
    np.testing.assert_array_equal(got.T, want)
    np.testing.assert_allclose(got[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=1e-15)
    # conventions the writer copies from this file: the heap's free list ends with 1, not with the undefined address
    (heap,) = f.group_info["heaps"].values()
    free = f._abs(heap["data_addr"]) + heap["free_head"]
    assert f._u(free, 8) == 1 and heap["free_head"] + f._u(free + 8, 8) == heap["data_size"]


def test_datatype_message_is_byte_identical_to_libhdf5(vpm, tmp_path):
    """the IEEE-754 little-endian binary64 datatype message of our datasets equals, byte for byte, the one the real
    library wrote for the float64 dataset of the sample file"""
    import scipy.io
    sample = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(sample):
        pytest.skip("scipy's MATLAB sample file is not installed")
    want = h5mini.File(sample).datasets["testdouble"].payloads[0x03]
    path = tmp_path / "dt.h5"
    vpm.H5Writer(path).create_dataset("z", (2, 5, 3)).commit().close()
    ours = h5mini.File(path).datasets["z"]
    assert ours.payloads[0x03] == want and len(want) == 24
    # fill value: the library's "default value" encoding (defined, size 0), here in message version 2
    assert ours.payloads[0x05] == bytes([2, 1, 2, 1, 0, 0, 0, 0])


def test_splitting_layout_roundtrip(vpm, tmp_path):
    nd, npart, nt = 2, 7, 5
    rng = np.random.default_rng(1)
    z = rng.standard_normal((nt + 1, npart, nd))            # file order = Julia (nd, np, nt+1) reversed
    path = tmp_path / "vp.h5"
    with vpm.H5Writer(path).create_dataset("z", (nd, npart, nt + 1)).create_dataset("t", (nt + 1,)).commit() as w:
        for n in (3, 0, 5, 1, 4, 2):                        # any order
            w.write_frame("z", n, z[n])
            w.write_frame("t", n, [0.1 * n])
    f = h5mini.File(path)
    assert f.eof_addr == os.path.getsize(path) and f.base == 0 and f.flags == 0
    assert sorted(f.datasets) == ["t", "z"]
    dz = f.datasets["z"]
    assert dz.shape == (nt + 1, npart, nd) and dz.maxshape == (None, npart, nd) and dz.chunk == (1, npart, nd)
    assert dz.dtype == np.dtype("<f8") and dz.layout == "chunked" and not dz.filters
    np.testing.assert_array_equal(f.read("z"), z)
    # what h5read(h5file, "z") returns in Julia: z[:, :, n+1] is the 2 x np state after step n
    zj = np.transpose(f.read("z"), (2, 1, 0))
    np.testing.assert_array_equal(zj[:, :, 3], z[3].T)
    dt = f.datasets["t"]
    assert dt.shape == (nt + 1,) and dt.maxshape == (None,) and dt.chunk == (1,)
    np.testing.assert_array_equal(f.read("t"), 0.1 * np.arange(nt + 1))


def test_geometric_integrator_layout_many_frames(vpm, tmp_path):
    """z[np, nt+1] chunk (np, 1) and t[nt+1] chunk (1): 5001 frames need a three-level chunk B-tree
    (64 entries per node), written in pieces as the device-to-host ring does"""
    npart, nframes = 33, 5001
    rng = np.random.default_rng(2)
    z = rng.standard_normal((nframes, npart))
    t = 0.01 * np.arange(nframes)
    path = tmp_path / "lb.h5"
    with vpm.H5Writer(path).create_dataset("z", (npart, nframes)).create_dataset("t", (nframes,)).commit() as w:
        for n in range(nframes):
            w.write_frame("z", n, z[n, :20])
            w.write_frame("z", n, z[n, 20:], offset=20)
            w.write_frame("t", n, t[n:n + 1])
    f = h5mini.File(path)
    np.testing.assert_array_equal(f.read("z"), z)
    np.testing.assert_array_equal(f.read("t"), t)
    nodes = f.group_info["chunk_nodes"]
    for ds in ("z", "t"):
        assert len(f.datasets[ds].chunks) == nframes
    assert max(n["level"] for n in nodes) == 2
    leaves = [n for n in nodes if n["level"] == 0]
    assert len(leaves) == 2 * 79 and all(n["used"] <= 64 for n in nodes)
    # sibling links form one left-to-right chain per level and dataset
    by_addr = {n["addr"]: n for n in nodes}
    for n in nodes:
        if n["right"] != h5mini.UNDEF:
            r = by_addr[n["right"]]
            assert r["left"] == n["addr"] and r["level"] == n["level"] and r["first"][2] == n["last"][2]
    assert f.eof_addr == os.path.getsize(path)


def test_unwritten_frames_read_as_zeros_and_file_is_valid_before_close(vpm, tmp_path):
    path = tmp_path / "partial.h5"
    w = vpm.H5Writer(path).create_dataset("z", (2, 3, 4)).commit()
    w.write_frame("z", 1, np.ones((3, 2)))
    got = h5mini.File(path).read("z")           # read while still open: metadata is complete after commit
    assert got.shape == (4, 3, 2) and got[1].min() == 1.0 and np.abs(got[[0, 2, 3]]).max() == 0.0
    w.close()
    w.close()                                   # idempotent


def test_writer_errors(vpm, tmp_path):
    with pytest.raises(vpm.VpmError, match="cannot open"):
        vpm.H5Writer(tmp_path / "no_such_dir" / "x.h5")
    w = vpm.H5Writer(tmp_path / "e.h5").create_dataset("z", (2, 4, 3))
    with pytest.raises(vpm.VpmError, match="duplicate"):
        w.create_dataset("z", (3,))
    with pytest.raises(vpm.VpmError, match="4 GiB"):
        w.create_dataset("big", (2, 300_000_000, 3))       # 4.8 GB per frame: HDF5's chunk-size limit
    with pytest.raises(vpm.VpmError, match="commit"):
        w.write_frame("z", 0, np.zeros(8))
    w.commit()
    with pytest.raises(vpm.VpmError, match="already committed"):
        w.create_dataset("t", (3,))
    for frame, data, off in ((3, np.zeros(8), 0), (-1, np.zeros(8), 0), (0, np.zeros(9), 0), (0, np.zeros(4), 5)):
        with pytest.raises(vpm.VpmError, match="outside"):
            w.write_frame("z", frame, data, offset=off)
    w.close()
    with pytest.raises(vpm.VpmError, match="no datasets"):
        vpm.H5Writer(tmp_path / "empty.h5").commit()


def test_writer_fuzz(vpm, tmp_path):
    """random files: 1-8 datasets of rank 1-3, 1-200 frames (one- and two-level chunk B-trees), random names, frames
    written in random order and random pieces; every dataset must read back exactly"""
    from hypothesis import given, settings, strategies as st

    names = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCXYZ_0123456789", min_size=1, max_size=12)
    dset = st.tuples(names, st.lists(st.integers(1, 6), min_size=0, max_size=2))

    @settings(max_examples=40, deadline=None)
    @given(st.lists(dset, min_size=1, max_size=8, unique_by=lambda d: d[0]), st.integers(1, 200), st.integers(0, 2**32 - 1))
    def run(dsets, nframes, seed):
        rng = np.random.default_rng(seed)
        path = tmp_path / "fuzz.h5"
        w = vpm.H5Writer(path)
        data = {}
        for name, inner in dsets:
            shape = tuple(inner) + (nframes,)                       # Julia order: frame axis last
            w.create_dataset(name, shape)
            data[name] = rng.standard_normal((nframes,) + tuple(reversed(inner)))
        w.commit()
        for name, a in data.items():
            for n in rng.permutation(nframes):
                flat = a[n].ravel()
                cut = int(rng.integers(0, flat.size + 1))
                if cut:
                    w.write_frame(name, n, flat[:cut])
                if cut < flat.size:
                    w.write_frame(name, n, flat[cut:], offset=cut)
        w.close()
        f = h5mini.File(path)
        assert sorted(f.datasets) == sorted(data) and f.eof_addr == os.path.getsize(path)
        for name, a in data.items():
            ds = f.datasets[name]
            assert ds.shape == a.shape and ds.chunk == (1,) + a.shape[1:] and ds.maxshape == (None,) + a.shape[1:]
            np.testing.assert_array_equal(f.read(name), a)

    run()


def test_mapped_multithreaded_writes(vpm, tmp_path, monkeypatch):
    """opt-in VPM_H5_THREADS > 1: the file is mapped and large pieces are copied by several threads; same bytes"""
    monkeypatch.setenv("VPM_H5_THREADS", "4")
    npart, nframes = 1_500_001, 3                           # 24 MB frames: pieces above and below the 1 MiB-per-thread cut
    rng = np.random.default_rng(3)
    z = rng.standard_normal((nframes, npart, 2))
    path = tmp_path / "mapped.h5"
    with vpm.H5Writer(path).create_dataset("z", (2, npart, nframes)).create_dataset("t", (nframes,)).commit() as w:
        for n in range(nframes):
            flat = z[n].ravel()
            cuts = [0, 7, 7 + (5 << 17), 7 + (5 << 17) + (3 << 18) + 1, flat.size]
            for a, b in zip(cuts[:-1], cuts[1:]):
                w.write_frame("z", n, flat[a:b], offset=a)
            w.write_frame("t", n, [float(n)])
    f = h5mini.File(path)
    np.testing.assert_array_equal(f.read("z"), z)
    np.testing.assert_array_equal(f.read("t"), np.arange(nframes, dtype=float))
    assert f.eof_addr == os.path.getsize(path)


def test_four_level_chunk_tree(vpm, tmp_path):
    """more than 64^3 frames (e.g. the reference's every-step output of a very long run): a four-level chunk B-tree"""
    n = 270_000
    path = tmp_path / "deep.h5"
    w = vpm.H5Writer(path).create_dataset("t", (n,)).commit()
    idx = [0, 1, 63, 64, 65, 4095, 4096, 4097, 262143, 262144, 262145, n - 1]     # both sides of every node boundary
    for i in idx:
        w.write_frame("t", i, [i + 0.5])
    w.close()
    f = h5mini.File(path)
    a = f.read("t")
    assert max(x["level"] for x in f.group_info["chunk_nodes"]) == 3 and len(f.datasets["t"].chunks) == n
    assert all(a[i] == i + 0.5 for i in idx) and a.sum() == sum(i + 0.5 for i in idx)


# ---------------------------------------------------------------------------------- structural validation
def _sample():
    import scipy.io
    return os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")


def test_validator_accepts_a_file_written_by_libhdf5():
    """the structural validator (tests/h5validate.py) is pinned on the real library's output first"""
    import h5validate
    if not os.path.exists(_sample()):
        pytest.skip("scipy's MATLAB sample file is not installed")
    rep = h5validate.validate(_sample())
    assert rep["base"] == 512 and rep["eof"] == rep["file_size"]
    assert rep["datasets"]["testdouble"]["layout"] == "contiguous" and rep["snods"] >= 1 and rep["allocations"] >= 6
    assert all(h["free_bytes"] + 8 <= h["data_size"] for h in rep["heaps"])


@pytest.mark.parametrize("nframes", [1, 6, 64, 65, 130, 5001])
def test_written_files_pass_structural_validation(vpm, tmp_path, nframes):
    """every address, size and key of our files: inside the file, disjoint, B-tree invariants (1, 2 and 3 levels), heap
    free list, symbol-table caches, superblock EOF == file size -- what H5Fopen / H5Dread check before touching data"""
    import h5validate
    path = tmp_path / "v.h5"
    npart = 3
    w = vpm.H5Writer(path).create_dataset("z", (2, npart, nframes)).create_dataset("t", (nframes,)).commit()
    for n in {0, nframes // 2, nframes - 1}:
        w.write_frame("z", n, np.full((npart, 2), float(n)))
        w.write_frame("t", n, [0.5 * n])
    w.close()
    rep = h5validate.validate(path)
    dz, dt = rep["datasets"]["z"], rep["datasets"]["t"]
    assert dz["chunks"] == dz["chunks_possible"] == nframes and dt["chunks"] == nframes
    assert dz["btree_depth"] == (1 if nframes <= 64 else 2 if nframes <= 64 * 64 else 3)
    assert rep["eof"] == rep["file_size"] == os.path.getsize(path)
    # the validated structure holds the data the reader returns
    f = h5mini.File(path)
    assert f.read("t")[nframes - 1] == 0.5 * (nframes - 1) and f.read("z")[nframes // 2, 0, 0] == float(nframes // 2)


def test_validator_rejects_corrupted_structures(vpm, tmp_path):
    """the validator is not vacuous: single-field corruptions of a good file are caught"""
    import h5validate
    path = tmp_path / "good.h5"
    w = vpm.H5Writer(path).create_dataset("z", (2, 3, 130)).create_dataset("t", (130,)).commit()
    w.close()
    good = open(path, "rb").read()
    h5validate.validate(path)
    f = h5mini.File(path)
    f.read("z")
    nodes = f.group_info["chunk_nodes"]
    leaf = next(n for n in nodes if n["level"] == 0)
    heap_addr, heap = next(iter(f.group_info["heaps"].items()))

    def corrupt(off, data):
        b = bytearray(good)
        b[off:off + len(data)] = data
        p = tmp_path / "bad.h5"
        p.write_bytes(bytes(b))
        return p

    eof_field = 24 + 16                                      # superblock v0: base, free-space, EOF, driver addresses
    cases = {
        "eof beyond file": corrupt(eof_field, (len(good) + 8).to_bytes(8, "little")),
        "eof short of file": corrupt(eof_field, (len(good) - 8).to_bytes(8, "little")),
        "btree key order": corrupt(leaf["addr"] + 24 + 8, (10 ** 6).to_bytes(8, "little")),
        "btree sibling": corrupt(leaf["addr"] + 16, (1234).to_bytes(8, "little")),
        "chunk size field": corrupt(leaf["addr"] + 24, (7).to_bytes(4, "little")),
        "heap free list": corrupt(heap_addr + 16, (heap["data_size"] + 64).to_bytes(8, "little")),
        "heap free terminator": corrupt(heap["data_addr"] + heap["free_head"], (0xFFFFFFFFFFFFFFFF).to_bytes(8, "little")),
        "object header count": corrupt(f.root_entry["ohdr"] + 2, (9).to_bytes(2, "little")),
    }
    for name, p in cases.items():
        with pytest.raises((h5mini.H5FormatError, ValueError, IndexError, struct_error())):
            h5validate.validate(p)
            pytest.fail(f"corruption not detected: {name}")


def struct_error():
    import struct
    return struct.error
