"""Shared by the script mirrors: repository import path and an `h5read` for the post-processing lines.

The reference scripts read their trajectory back with HDF5.jl (`z = h5read(h5file, "z")`).  This image has neither
HDF5.jl nor h5py, so `h5read` uses h5py when it is importable and otherwise the small independent reader that the
test-suite uses to check the library's HDF5 writer (tests/h5mini.py).  Arrays are returned with Julia's dimension
order (frame axis last), so the post-processing reads like upstream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def h5read(path, name):
    import numpy as np
    try:
        import h5py
        with h5py.File(path, "r") as f:
            a = f[name][...]
    except ImportError:
        import h5mini
        a = h5mini.File(path).read(name)
    return np.transpose(a)          # HDF5 row-major (nt+1, np, nd) -> Julia (nd, np, nt+1)
