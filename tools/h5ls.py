#!/usr/bin/env python
"""List the datasets of a trajectory file written by run!(method, h5file) (no HDF5 library needed; uses the
independent reader of the test-suite, tests/h5mini.py).

    python tools/h5ls.py run.h5            # name, shape (HDF5 order and Julia order), chunks, first / last frame summary
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(path):
    import numpy as np
    import h5mini
    f = h5mini.File(path)
    print(f"{path}: superblock v{f.sb_version}, {os.path.getsize(path)} bytes, end-of-file address {f.eof_addr}")
    for name, ds in sorted(f.datasets.items()):
        julia = tuple(reversed(ds.shape))
        print(f"  {name}: {ds.dtype} shape {ds.shape} (Julia {julia}), max {ds.maxshape}, {ds.layout}"
              + (f" chunks {ds.chunk}" if ds.chunk else ""))
        if np.prod(ds.shape) * ds.dtype.itemsize <= 2e9:
            a = f.read(name)
            first, last = (a[0], a[-1]) if a.ndim > 1 else (a[:1], a[-1:])
            print(f"    first frame: min {np.min(first):.6g} max {np.max(first):.6g}; last frame: min {np.min(last):.6g} max {np.max(last):.6g}")


if __name__ == "__main__":
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    main(sys.argv[1])
