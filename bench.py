#!/usr/bin/env python
"""bench.py — particle-steps/s of the fused Vlasov-Poisson Strang step on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: oracle port of the reference algorithm

Workload (config.workload): BASELINE.json configs[1], scripts/bump_on_tail.jl at 1e8 particles per GPU:
L = 2 pi / 0.3, n_h = 16 periodic splines of degree 3 (order 4), dt = 0.1, chi = 1, self-consistent Strang
(legacy integrate_vp!, src/vlasov_poisson.jl:94-115).  A "step" is one Strang step of every particle.
N > 1 (torchrun, one rank per GPU) is BASELINE configs[4]: the same physics with 1e8 particles per GPU
(weak scaling), particle slabs per rank and one all-reduce of the 16+2 coefficient vector per step.

Timed region: ONE call of the whole-step stepper for K steps (K+1 streaming passes: the half-drift
staggering costs one extra pass per call), CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  Inputs (2.4 GB/GPU) are far larger than L2, so no flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

KAPPA, NH, ORDER, DT, CHI = 0.3, 16, 4, 0.1, 1.0
L = 2 * np.pi / KAPPA
BYTES_PER_STEP = 40  # read x,v,w + write x,v (fp64 SoA), SURVEY 8d / BASELINE.md section 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=1e8, help="particles per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=float, default=1e7, help="particles of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="vp", choices=["vp", "lb", "clb"])
    ap.add_argument("--load", default="bump_on_tail", choices=["bump_on_tail", "uniform", "one_cell", "sorted"],
                    help="vp particle load: the config-2 sampler (default) or the synthetic loads of SURVEY 8(d): uniform x / "
                         "all particles in one cell (worst case for atomics-based deposits) / cell-sorted x")
    ap.add_argument("--n-basis", type=int, default=16, help="x-space basis size (headline config: 16)")
    ap.add_argument("--order", type=int, default=4, help="spline order (headline config: 4 = cubic)")
    ap.add_argument("--field", default="selfconsistent", choices=["selfconsistent", "frozen"],
                    help="vp: self-consistent Strang loop (headline) or the frozen field of the shipped SplittingMethod")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: fused peer-memory all-reduce in the field kernel (default) or NCCL")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, c in zip(names, r[5:9]) if c.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_rate(nsample, nsteps, threads):
    """particle-steps/s of the oracle port (faithful restatement of the reference loops) on the host cores."""
    from oracle import oracle as orc
    orc.set_threads(threads)
    x, v, w = orc.sample_bump_on_tail(int(nsample), kappa=KAPPA)
    xs = orc.XSpace(0.0, L, ORDER, NH)
    t0 = time.perf_counter()
    xs.strang_selfconsistent(x, v, w, DT, nsteps, chi=CHI, diag=False)
    dt = time.perf_counter() - t0
    orc.set_threads(1)
    return nsample * nsteps / dt, dt


def run_reference(args):
    """CPU arm: the reference is pure Julia and cannot run here (no Julia toolchain), so this
    times the oracle port of its algorithm with all host threads on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = orc.max_threads()
    nsample = int(args.cpu_sample)
    for _ in range(min(args.warmup, 1)):
        cpu_port_rate(nsample // 4, 1, cores)
    rate, dt = cpu_port_rate(nsample, args.steps if args.steps <= 20 else 20, cores)
    ksteps = args.steps if args.steps <= 20 else 20
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": rate, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": ksteps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / ksteps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "vp_bump_on_tail_strang_selfconsistent", "particles_per_step": nsample, "n_basis": NH,
                   "order": ORDER, "dt": DT, "note": "bounded sample of the 1e8-particle workload; throughput per particle-step"},
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{nsample} particles x {ksteps} Strang steps, OpenMP {cores} threads, oracle C port (reference is Julia: not runnable here)"},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    global NH, ORDER
    args = parse()
    NH, ORDER = args.n_basis, args.order
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import vpm_b200 as vpm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvpm_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a dedicated non-default stream shared by torch (events) and the library (kernels): the legacy default
    # stream has handle 0, which the C ABI reads as "create a private stream"
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = vpm.Context(local, stream.cuda_stream)
    vpm.set_default_context(ctx)
    comm_used = "none"
    if world > 1:
        if args.comm == "p2p":
            # fused peer-memory all-reduce inside the field kernel (p2p.cuh); NCCL is the fallback
            try:
                handles = [None] * world
                dist.all_gather_object(handles, ctx.p2p_prepare())
                ctx.p2p_attach(world, rank, handles)
                ok = torch.tensor([1], device="cuda")
            except vpm.VpmError as e:
                print(f"[rank {rank}] p2p unavailable ({e}); falling back to NCCL", file=sys.stderr)
                ok = torch.tensor([0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                comm_used = "p2p"
            else:
                ctx.p2p_detach()
        if comm_used != "p2p":
            obj = [vpm.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            ctx.comm_init(world, rank, obj[0])
            comm_used = "nccl"

    n = int(args.particles)
    ntotal = n * world
    lib = vpm._cabi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "vp":
        d = vpm.ParticleDistribution(1, 1, n, ctx)
        vpm.initialize_(d, vpm.BumpOnTail(kappa=KAPPA), offset=rank * n, ntotal=ntotal)
        if args.load != "bump_on_tail":   # synthetic throughput loads, generated on the host (seeded), w = L / N
            rng = np.random.default_rng(0x5EED0002 + rank)
            if args.load == "one_cell":    # every particle in cell 5 and slow enough to stay there for the whole run
                xs_, vs_ = (5.3 + 0.2 * rng.random(n)) * (L / NH), 1e-4 * rng.standard_normal(n)
            else:
                xs_, vs_ = L * rng.random(n), rng.standard_normal(n)
                if args.load == "sorted":
                    xs_.sort()
            d.set(xs_, vs_, np.full(n, L / ntotal))
            del xs_, vs_
        pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), ORDER, NH), ctx)

        def run_steps(k):
            vpm.check(lib.vpm_vp_strang_steps_async(pot._h, d._h, DT, CHI, int(k), 1 if args.field == "frozen" else 0, 0))
        kind_pass, bytes_unit, wl = 0, BYTES_PER_STEP, "vp_" + args.load + "_strang_" + args.field
        passes_per_call = lambda k: k + 1
    else:
        cons = args.workload == "clb"
        d = vpm.ParticleDistribution(1, 1, n, ctx)
        vpm.initialize_(d, vpm.DoubleMaxwellian((-10.0, 10.0), 2.0), offset=rank * n, ntotal=ntotal)
        sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet", ctx)

        def run_steps(k):
            vpm.check(lib.vpm_lb_rk438_steps_async(sd._h, d._h, 1.0, 1e-2, int(k), int(cons)))
        # RK438 particle-step in k form: stage passes 24+32+48+32 B (CLB: +8 B q stores in stages 1, 2 and 4 x 8 B moment passes), DESIGN.md
        kind_pass, bytes_unit, wl = 2, (136 + (48 if cons else 0)), ("clb" if cons else "lb") + "_rk438_double_maxwellian"
        passes_per_call = lambda k: 4 * k * (2 if cons else 1) + 1

    # ---- warm-up, then the timed region ----
    run_steps(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = ntotal * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream ----
    vpm.check(lib.vpm_profile(ctx._h, 1))
    ksteps = min(args.steps, 20)
    run_steps(ksteps)
    msk = (np.zeros(8), np.zeros(8, dtype=np.int64))
    vpm.check(lib.vpm_profile_get(ctx._h, msk[0].ctypes.data, msk[1].ctypes.data))
    vpm.check(lib.vpm_profile(ctx._h, 0))
    pass_ms, pass_cnt = float(msk[0][kind_pass]), int(msk[1][kind_pass])
    field_ms = float(msk[0][kind_pass + 1])
    if args.workload == "vp":
        bytes_per_launch = BYTES_PER_STEP * n           # one fused pass = one particle-step of every particle
    else:
        bytes_per_launch = bytes_unit * n * ksteps / max(pass_cnt, 1)   # mean over the stage / moment passes
    achieved = bytes_per_launch / (pass_ms / max(pass_cnt, 1) * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)

    # ---- context only (NOT the headline): uniform-weight fast path, 32 B per particle-step -------------------
    # every reference sampler produces equal weights (w = L/N), so the steppers can skip the w[] stream when the
    # caller declares it; `value` above always streams w[] (40 B), as the reference's layout does.
    uniform = None
    if args.workload == "vp" and world == 1:
        d.set_uniform_weight(L / ntotal)
        run_steps(3)
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ku = min(args.steps, 50)
        u0.record(stream)
        run_steps(ku)
        u1.record(stream)
        barrier()
        ums = u0.elapsed_time(u1) / ku
        uniform = {"value": n / (ums * 1e-3), "unit": "particle-steps/s", "ms_per_step": ums,
                   "bytes_per_particle_step": 32, "GBps": 32 * n / (ums * 1e-3) / 1e9,
                   "api": "vpm_particles_set_uniform_weight + vpm_vp_strang_steps"}
        vpm.initialize_(d, vpm.BumpOnTail(kappa=KAPPA), offset=rank * n, ntotal=ntotal)   # back to per-particle weights

    # ---- e2e: the host-array drop-in step (z = 2 x N host matrix in, out), PCIe copies inside the timing ----
    e2e = None
    if args.workload == "vp" and not args.no_e2e:
        import ctypes as C
        zin, zout = C.c_void_p(), C.c_void_p()
        vpm.check(lib.vpm_host_alloc(16 * n, C.byref(zin)))
        vpm.check(lib.vpm_host_alloc(16 * n, C.byref(zout)))
        vpm.check(lib.vpm_particles_download_aos(d._h, zin, 2))
        vpm.check(lib.vpm_vp_strang_step_host(pot._h, d._h, zin, zout, DT, CHI, 0))  # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        cur, nxt = zin, zout
        for _ in range(args.e2e_steps):
            vpm.check(lib.vpm_vp_strang_step_host(pot._h, d._h, cur, nxt, DT, CHI, 0))
            cur, nxt = nxt, cur
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": ntotal * args.e2e_steps / float(t_e2e.item()), "unit": "particle-steps/s",
               "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n, "steps": args.e2e_steps,
               "api": "vpm_vp_strang_step_host (pinned 2 x N integrator state in/out per step; weights resident)"}
        # for context: what a run!(integrator) user gets — initial state uploaded once from pinned memory, K steps
        # on the device with (W,K,M) read back for every step, final state downloaded once
        kk = min(args.steps, 100)
        diag = np.zeros((kk + 1, 3))
        barrier()
        t0 = time.perf_counter()
        vpm.check(lib.vpm_particles_upload_aos(d._h, cur, 2))
        vpm.check(lib.vpm_vp_strang_steps(pot._h, d._h, DT, CHI, kk, 0, 1, diag.ctypes.data))
        vpm.check(lib.vpm_particles_download_aos(d._h, nxt, 2))
        barrier()
        t_run = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_run, op=dist.ReduceOp.MAX)
        e2e["run_api"] = {"value": ntotal * kk / float(t_run.item()), "unit": "particle-steps/s", "steps": kk,
                          "h2d_bytes_total": 16 * n, "d2h_bytes_total": 16 * n + 24 * (kk + 1),
                          "api": "upload_aos + vpm_vp_strang_steps(diag_mode=1) + download_aos (state resident between steps)"}
        lib.vpm_host_free(zin)
        lib.vpm_host_free(zout)

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.workload == "vp":
        from oracle import oracle as orc
        cores = orc.max_threads()
        r1, _ = cpu_port_rate(int(args.cpu_sample) // 4, 2, 1)
        rall, _ = cpu_port_rate(int(args.cpu_sample), 10, cores)
        cpu = {"value": rall, "unit": "particle-steps/s", "cores": cores, "kind": "port",
               "sample": f"{int(args.cpu_sample)} particles x 10 Strang steps on {cores} OpenMP threads; "
                         f"1 thread ({int(args.cpu_sample) // 4} x 2): {r1:.3e}/s; C port of the reference algorithm (Julia not runnable here)",
               "single_thread_value": r1}

    if rank == 0:
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "particles_per_gpu": n, "n_basis": NH if args.workload == "vp" else 41,
                       "order": ORDER, "dt": DT if args.workload == "vp" else 1e-2,
                       "l2": "inputs (24 B x particles per GPU) larger than L2, no flush needed",
                       "parallelism": f"particle slabs x{world}, coefficient all-reduce per field update ({comm_used})" if world > 1 else "single GPU",
                       "passes_in_timed_region": passes_per_call(args.steps)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": ("vp_pass_tma_kernel (fused kick+drift+deposit; prologue/epilogue passes use vp_pass_kernel)"
                                    if os.environ.get("VPM_TUNE_TMA", "1") != "0" else "vp_pass_kernel") if args.workload == "vp" else "lb_pass_kernel",
                         "bytes_per_launch": bytes_per_launch, "avg_launch_ms": pass_ms / max(pass_cnt, 1),
                         "launches_timed": pass_cnt, "field_kernel_share": field_ms / max(pass_ms + field_ms, 1e-30),
                         "frac_of_8TBs_nominal": achieved / 8000.0},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "uniform_weight_variant": uniform,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        if comm_used == "p2p":
            ctx.p2p_check()
            dist.barrier()
            ctx.p2p_detach()
        else:
            ctx.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
