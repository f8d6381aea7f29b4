"""world_size-2 gloo test of the multi-GPU host logic: particle slabs are generated from
(seed, global index), every rank deposits its slab, the coefficient vector (+K, M) is all-reduced and the
solve is replicated.  The kernels are replaced by the oracle here (tests may use it); on the GPU the same
plumbing runs with libvpm_b200 + NCCL (bench.py --gpus N, tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nper, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, nh, L = 4, 16, 2 * np.pi / 0.3
    xs = orc.XSpace(0.0, L, K, nh)
    # the 128-byte communicator id is created on rank 0 and broadcast as an object (as bench.py does)
    obj = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == bytes(range(128))
    x, v, w = orc.sample_bump_on_tail(nper, offset=rank * nper, Ntotal=world * nper)
    x = orc.push_drift(x, v, 0.05)
    buf = np.zeros(nh + 2)
    buf[:nh] = xs.deposit(x, w)
    buf[nh] = (w * v * v).sum()
    buf[nh + 1] = (w * v).sum()
    t = torch.from_numpy(buf)
    dist.all_reduce(t)                      # the one collective of the path: nh + 2 doubles
    phi = xs.poisson_solve(buf[:nh])        # replicated solve on bitwise-identical input
    v1 = xs.push_kick(phi, x, v, 0.1)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), buf=buf, phi=phi, x=x, v1=v1)
    dist.destroy_process_group()


def test_slab_sharding_matches_single_rank(tmp_path, oracle):
    world, nper = 2, 6000
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nper, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # every rank holds the same reduced vector and therefore the same field, bit for bit
    np.testing.assert_array_equal(r[0]["buf"], r[1]["buf"])
    np.testing.assert_array_equal(r[0]["phi"], r[1]["phi"])
    # ... and it equals the single-rank computation up to summation order
    K, nh, L = 4, 16, 2 * np.pi / 0.3
    xs = oracle.XSpace(0.0, L, K, nh)
    x, v, w = oracle.sample_bump_on_tail(world * nper)
    x = oracle.push_drift(x, v, 0.05)
    rhs = xs.deposit(x, w)
    assert np.linalg.norm(r[0]["buf"][:nh] - rhs) < 1e-13 * np.linalg.norm(rhs)
    phi = xs.poisson_solve(rhs)
    assert np.linalg.norm(r[0]["phi"] - phi) < 1e-11 * np.linalg.norm(phi)
    v1 = xs.push_kick(phi, x, v, 0.1)
    got = np.concatenate([r[0]["v1"], r[1]["v1"]])
    assert np.linalg.norm(got - v1) < 1e-12 * np.linalg.norm(v1)
    np.testing.assert_array_equal(np.concatenate([r[0]["x"], r[1]["x"]]), x)
