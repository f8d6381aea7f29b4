#!/usr/bin/env python
"""Host-only micro-benchmark of the trajectory writer (no GPU): GB/s of vpm_h5_write into a fresh file, written in
32 MiB pieces as the device-to-host ring does, with the default pwrite path and with the opt-in mapped multi-threaded
path (VPM_H5_THREADS), on tmpfs and on the temp directory's file system.  Prints one JSON object.

    python tools/h5_write_bench.py [--mb 2048] [--threads 1,4,8]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(path, mb):
    import numpy as np
    import vpm_b200 as vpm
    piece = np.ones(4 << 20)                  # 32 MiB
    npieces = max(mb // 32, 1)
    nframes = 2
    per_frame = npieces // nframes or 1
    w = vpm.H5Writer(path).create_dataset("z", (per_frame * piece.size, nframes)).commit()
    t0 = time.perf_counter()
    for n in range(nframes):
        for i in range(per_frame):
            w.write_frame("z", n, piece, offset=i * piece.size)
    dt = time.perf_counter() - t0
    w.close()
    os.remove(path)
    print(json.dumps({"GBps": nframes * per_frame * piece.nbytes / dt / 1e9}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--threads", default="1,4,8")
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child:
        return child(a.child, a.mb)
    out = {"mb": a.mb, "cores": os.cpu_count()}
    dirs = [d for d in ("/dev/shm", tempfile.gettempdir()) if os.path.isdir(d)]
    for d in dirs:
        for t in [int(x) for x in a.threads.split(",")]:
            env = dict(os.environ, VPM_H5_THREADS=str(t))
            r = subprocess.run([sys.executable, __file__, "--child", os.path.join(d, f"vpm_h5_bench_{os.getpid()}.h5"), "--mb", str(a.mb)],
                               env=env, capture_output=True, text=True)
            out[f"{d} threads={t}" + (" (pwrite)" if t == 1 else " (mmap)")] = json.loads(r.stdout.strip().splitlines()[-1])["GBps"] if r.returncode == 0 else r.stderr[-300:]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
