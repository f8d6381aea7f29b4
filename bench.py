#!/usr/bin/env python
"""bench.py — particle-steps/s of the fused Vlasov-Poisson Strang step on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: oracle port of the reference algorithm

Workload (config.workload): BASELINE.json configs[1], scripts/bump_on_tail.jl at 1e8 particles per GPU:
L = 2 pi / 0.3, n_h = 16 periodic splines of degree 3 (order 4), dt = 0.1, chi = 1, self-consistent Strang
(legacy integrate_vp!, src/vlasov_poisson.jl:94-115).  A "step" is one Strang step of every particle.
N > 1 (torchrun, one rank per GPU) is BASELINE configs[4]: the same physics with 1e8 particles per GPU
(weak scaling), particle slabs per rank and one all-reduce of the 16+2 coefficient vector per step.

Timed region of `value`: ONE call of the whole-step stepper for K steps (K+1 streaming passes: the half-drift
staggering costs one extra pass per call), CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  Inputs (2.4 GB/GPU) are far larger than L2, so no flush is needed.

The default line also carries (all measured in the same run, none inside the timed region of `value`):
  sustained      the same stepper for >= 5 s with the clocks / power sampled during it
  workloads      BASELINE configs[2], [3]: Lenard-Bernstein and conservative LB RK438 steps (1e8 particles/GPU),
                 per-pass roofline, sustained window, CPU port beside them
  e2e            the host-array drop-in step (PCIe inside the timing) and the box's measured PCIe ceiling
  parity_check   a small run of VP + CLB on the same ranks / communicator, checked against a single-rank run and
                 the CPU oracle (exit code 3 if it fails)
"""
import argparse
import ctypes as C
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
LB_NKNOTS, LB_ORDER, LB_DOMAIN, LB_NU, LB_DT, LB_SHIFT = 41, 4, (-10.0, 10.0), 1.0, 1e-2, 2.0
# RK438 particle-step in k form: bytes per particle of each pass kind (vpm_profile_get_lb index)
LB_PASS_BYTES = {0: 16, 1: 24, 2: 32, 3: 48, 4: 32, 6: 8}
LB_PASS_NAME = {0: "deposit_only", 1: "stage1", 2: "stage2", 3: "stage3", 4: "stage4", 6: "moments"}
PARITY_TOL = 1e-12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=1e8, help="particles per GPU")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--cpu-sample", type=float, default=1e7, help="particles of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / workloads / parity_check (tuning runs)")
    ap.add_argument("--sustained", action="store_true", help="with --no-extras: still run the sustained window")
    ap.add_argument("--sustained-seconds", type=float, default=5.5)
    ap.add_argument("--lb-steps", type=int, default=30, help="RK438 steps of the lb / clb blocks of the default line")
    ap.add_argument("--ref-budget-seconds", type=float, default=150.0,
                    help="--impl reference: bound of the timed CPU work; the particle count is reduced (and said so) beyond it")
    ap.add_argument("--workload", default="vp", choices=["vp", "lb", "clb"])
    ap.add_argument("--load", default="bump_on_tail", choices=["bump_on_tail", "uniform", "one_cell", "sorted"],
                    help="vp particle load: the config-2 sampler (default) or the synthetic loads of SURVEY 8(d): uniform x / "
                         "all particles in one cell (worst case for atomics-based deposits) / cell-sorted x")
    ap.add_argument("--n-basis", type=int, default=16, help="x-space basis size (headline config: 16)")
    ap.add_argument("--order", type=int, default=4, help="spline order (headline config: 4 = cubic)")
    ap.add_argument("--lb-nknots", type=int, default=41, help="lb / clb: knots of the v-grid (BASELINE configs: 41)")
    ap.add_argument("--field", default="selfconsistent", choices=["selfconsistent", "frozen"],
                    help="vp: self-consistent Strang loop (headline) or the frozen field of the shipped SplittingMethod")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: fused peer-memory all-reduce in the field kernel (default) or NCCL")
    return ap.parse_args()


def workload_config(args, world):
    """the `config` object: identical for both arms (the reference arm times the same workload on the host cores)"""
    n = int(args.particles)
    if args.workload == "vp":
        wl, nb, order, dt = "vp_" + args.load + "_strang_" + args.field, NH, ORDER, DT
    else:
        wl, nb, order, dt = args.workload + "_rk438_double_maxwellian", LB_NKNOTS, LB_ORDER, LB_DT
    return {"workload": wl, "particles_per_gpu": n, "n_basis": nb, "order": order, "dt": dt,
            "l2": "inputs (24 B x particles per GPU) larger than L2, no flush needed",
            "parallelism": f"particle slabs x{world}, coefficient all-reduce per field update" if world > 1 else "single GPU"}


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

        def num(r, i):
            try:
                return float(r[i])
            except (ValueError, IndexError):
                return None
        rows = [r for r in self.rows if len(r) >= 9]
        sm = [num(r, 1) for r in rows if num(r, 1) is not None]
        mx = [num(r, 2) for r in rows if num(r, 2) is not None]
        pw = [num(r, 3) for r in rows if num(r, 3) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, c in zip(names, r[5:9]) if c.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": min(sm) if sm else None, "power_w_median": float(np.median(pw)) if pw else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU port (oracle) legs: cpu_baseline of our line and the whole --impl reference arm
# ------------------------------------------------------------------------------------------------------------
def cpu_vp_rate(nsample, nsteps, threads, xvw=None):
    """particle-steps/s of the oracle port (faithful restatement of the reference loops) on the host cores."""
    from oracle import oracle as orc
    orc.set_threads(threads)
    x, v, w = xvw if xvw is not None else orc.sample_bump_on_tail(int(nsample), kappa=KAPPA)
    xs = orc.XSpace(0.0, L, ORDER, NH)
    t0 = time.perf_counter()
    xs.strang_selfconsistent(x, v, w, DT, nsteps, chi=CHI, diag=False)
    dt = time.perf_counter() - t0
    orc.set_threads(1)
    return len(x) * nsteps / dt, dt


def cpu_lb_rate(nsample, nsteps, threads, cons, vw=None):
    from oracle import oracle as orc
    orc.set_threads(threads)
    if vw is None:
        _, v, w = orc.sample_maxwellian(int(nsample), xlo=LB_DOMAIN[0], xhi=LB_DOMAIN[1], shift=LB_SHIFT, doubled=True)
    else:
        v, w = vw
    vs = orc.VSpace(LB_DOMAIN[0], LB_DOMAIN[1], LB_NKNOTS, LB_ORDER)
    t0 = time.perf_counter()
    vs.rk438(v, w, LB_NU, LB_DT, nsteps, conservative=cons, diag=False)
    dt = time.perf_counter() - t0
    orc.set_threads(1)
    return len(v) * nsteps / dt, dt


def cpu_baseline_block(kind, nsample, cores):
    """bounded sample of one workload on all host threads + one thread (the reference itself is serial,
    src/models/vlasov_poisson.jl:70-72)"""
    if kind == "vp":
        r1, _ = cpu_vp_rate(nsample // 4, 2, 1)
        rall, _ = cpu_vp_rate(nsample, 10, cores)
        what = f"{nsample} particles x 10 Strang steps on {cores} OpenMP threads; 1 thread ({nsample // 4} x 2): {r1:.3e}/s"
    else:
        ns = max(nsample // 5, 1000)
        r1, _ = cpu_lb_rate(ns // 4, 1, 1, kind == "clb")
        rall, _ = cpu_lb_rate(ns, 3, cores, kind == "clb")
        what = f"{ns} particles x 3 RK438 steps on {cores} OpenMP threads; 1 thread ({ns // 4} x 1): {r1:.3e}/s"
    return {"value": rall, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": what + "; C port of the reference algorithm (Julia not runnable here)", "single_thread_value": r1}


def run_reference(args):
    """CPU arm: the reference is pure Julia and cannot run here (no Julia toolchain), so this times the
    oracle port of its algorithm with all host threads ON THE SAME CONFIG: --particles, --steps and --warmup are
    honoured; only if the timed work would exceed --ref-budget-seconds is the particle count reduced (and reported)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = orc.max_threads()
    n = int(args.particles)
    kind = args.workload
    cons = kind == "clb"
    # calibrate on a small sample, then bound the work
    if kind == "vp":
        rate0, _ = cpu_vp_rate(1_000_000, 2, cores)
    else:
        rate0, _ = cpu_lb_rate(400_000, 1, cores, cons)
    nmax = int(args.ref_budget_seconds * rate0 / max(args.steps + args.warmup, 1))
    nrun = min(n, max(nmax, 100_000))
    orc.set_threads(cores)
    if kind == "vp":
        data = orc.sample_bump_on_tail(nrun, kappa=KAPPA)
        timed = lambda k: cpu_vp_rate(nrun, k, cores, data)
        single = cpu_vp_rate(min(nrun, 2_500_000), 2, 1)[0]
    else:
        _, v, w = orc.sample_maxwellian(nrun, xlo=LB_DOMAIN[0], xhi=LB_DOMAIN[1], shift=LB_SHIFT, doubled=True)
        timed = lambda k: cpu_lb_rate(nrun, k, cores, cons, (v, w))
        single = cpu_lb_rate(min(nrun, 500_000), 1, 1, cons)[0]
    if args.warmup > 0:
        timed(args.warmup)          # untimed warm-up steps on the same arrays (the stepper copies its inputs)
    rate, dt = timed(args.steps)
    sample = (f"{nrun} particles x {args.steps} steps (warm-up {args.warmup}) on {cores} OpenMP threads, oracle C port "
              f"(reference is Julia: not runnable here)")
    if nrun < n:
        sample += f"; particle count reduced from {n} to keep the timed CPU work under {args.ref_budget_seconds:.0f} s (throughput is per particle-step)"
    cfg = workload_config(args, args.gpus)
    if nrun < n:
        cfg["sample_particles"] = nrun
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": rate, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "single_thread_value": single},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if kind == "vp" and not args.no_extras:
        ns = int(args.cpu_sample)
        line["workloads"] = {k: {"value": b["value"], "unit": b["unit"], "cpu_baseline": b}
                             for k, b in ((k, cpu_baseline_block(k, ns, cores)) for k in ("lb", "clb"))}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import vpm_b200 as vpm
        self.args, self.torch, self.dist, self.vpm = args, torch, dist, vpm
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libvpm_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        # a dedicated non-default stream shared by torch (events) and the library (kernels): the legacy default
        # stream has handle 0, which the C ABI reads as "create a private stream"
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.ctx = vpm.Context(self.local, self.stream.cuda_stream)
        vpm.set_default_context(self.ctx)
        self.lib = vpm._cabi.lib()
        # host thread (and the pinned buffers it allocates from here on) onto the NUMA node of this rank's GPU
        self.numa = None
        import ctypes as C
        cpus = (C.c_char * 256)()
        if self.lib.vpm_ctx_bind_numa(self.ctx._h, cpus, 256) == 0:
            self.numa = cpus.value.decode()
        self.comm_used = "none"
        if self.world > 1:
            self._attach_comm()
        self.n = int(args.particles)
        self.ntotal = self.n * self.world
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def _attach_comm(self):
        vpm, dist, torch, ctx = self.vpm, self.dist, self.torch, self.ctx
        if self.args.comm == "p2p":
            # fused peer-memory all-reduce inside the field kernel (p2p.cuh); NCCL is the fallback
            try:
                handles = [None] * self.world
                dist.all_gather_object(handles, ctx.p2p_prepare())
                ctx.p2p_attach(self.world, self.rank, handles)
                ok = torch.tensor([1], device="cuda")
            except vpm.VpmError as e:
                print(f"[rank {self.rank}] p2p unavailable ({e}); falling back to NCCL", file=sys.stderr)
                ok = torch.tensor([0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                self.comm_used = "p2p"
            else:
                ctx.p2p_detach()
        if self.comm_used != "p2p":
            obj = [vpm.Context.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            ctx.comm_init(self.world, self.rank, obj[0])
            self.comm_used = "nccl"

    def barrier(self):
        if self.world > 1:
            if not hasattr(self, "_align"):
                self._align = self.torch.zeros(1, device="cuda")
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, run_steps, steps, sample_clocks=True):
        """K steps in one stepper call: events on the launching stream, barrier + sync both sides, max over ranks"""
        torch = self.torch
        sampler = ClockSampler(self.local)
        if self.rank == 0 and sample_clocks:
            sampler.start()
            time.sleep(0.3)
        l0 = self.ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if self.world > 1:
            # device-side rendezvous on the launching stream: the host-side barrier leaves the ranks' launch times ~100 us apart,
            # which the first fused all-reduce of the timed region would absorb (a fixed cost per call, visible in short windows)
            self.dist.all_reduce(self._align)
        e0.record(self.stream)
        run_steps(steps)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = self.ctx.launches - l0
        clocks = sampler.stop() if (self.rank == 0 and sample_clocks) else None
        return self.max_over_ranks(ms), int(launches), clocks

    def profile(self, run_steps, ksteps):
        """per-launch CUDA events recorded inside the library (vpm_profile): ms and counts by kernel kind and LB pass kind"""
        vpm, lib, ctx = self.vpm, self.lib, self.ctx
        vpm.check(lib.vpm_profile(ctx._h, 1))
        run_steps(ksteps)
        ms, cnt = np.zeros(8), np.zeros(8, dtype=np.int64)
        vpm.check(lib.vpm_profile_get(ctx._h, ms.ctypes.data, cnt.ctypes.data))
        msl, cntl = np.zeros(8), np.zeros(8, dtype=np.int64)
        vpm.check(lib.vpm_profile_get_lb(ctx._h, msl.ctypes.data, cntl.ctypes.data))
        vpm.check(lib.vpm_profile(ctx._h, 0))
        return ms, cnt, msl, cntl

    def sustained(self, run_steps, ms_per_step, bytes_per_step):
        """the same stepper for >= sustained_seconds: this pool's B200s drop their SM clock under sustained load
        (sw_power_cap), which an instruction-co-limited pass feels; clocks and power are sampled during the window"""
        steps = max(int(self.args.sustained_seconds * 1e3 / ms_per_step), 10)
        ms, _, clocks = self.timed(run_steps, steps)
        per = ms / steps
        gbs = bytes_per_step * self.n / (per * 1e-3) / 1e9
        return {"steps": steps, "seconds": ms * 1e-3, "ms_per_step": per, "value": self.ntotal / (per * 1e-3),
                "unit": "particle-steps/s", "whole_step_GBps": gbs, "frac": gbs / self.peak,
                "slowdown_vs_burst": per / ms_per_step, "clocks": clocks,
                "note": "frac = algorithmic bytes of the WHOLE step (passes + field kernels + gaps) / measured HBM peak"}

    # ---------------------------------------------------------------------------------------------------- VP
    def bench_vp(self):
        args, vpm, lib, ctx, n = self.args, self.vpm, self.lib, self.ctx, self.n
        d = vpm.ParticleDistribution(1, 1, n, ctx)
        vpm.initialize_(d, vpm.BumpOnTail(kappa=KAPPA), offset=self.rank * n, ntotal=self.ntotal)
        if args.load != "bump_on_tail":   # synthetic throughput loads, generated on the host (seeded), w = L / N
            rng = np.random.default_rng(0x5EED0002 + self.rank)
            if args.load == "one_cell":    # every particle in cell 5 and slow enough to stay there for the whole run
                xs_, vs_ = (5.3 + 0.2 * rng.random(n)) * (L / NH), 1e-4 * rng.standard_normal(n)
            else:
                xs_, vs_ = L * rng.random(n), rng.standard_normal(n)
                if args.load == "sorted":
                    xs_.sort()
            d.set(xs_, vs_, np.full(n, L / self.ntotal))
            del xs_, vs_
        pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), ORDER, NH), ctx)
        self.vp_d, self.vp_pot = d, pot
        mode = 1 if args.field == "frozen" else 0

        def run_steps(k):
            vpm.check(lib.vpm_vp_strang_steps_async(pot._h, d._h, DT, CHI, int(k), mode, 0))

        run_steps(max(args.warmup, 3))
        self.barrier()
        ms, launches, clocks = self.timed(run_steps, args.steps)
        per = ms / args.steps
        kms, kcnt, _, _ = self.profile(run_steps, min(args.steps, 20))
        pass_ms, pass_cnt, field_ms = float(kms[0]), int(kcnt[0]), float(kms[1])
        # one fused pass = one particle-step of every particle (40 B).  With the carried stagger (default, self-consistent mode)
        # a stepper call is exactly `steps` fused passes of 40 B; without it (VPM_TUNE_VPCARRY=0) the K + 1 passes of a call
        # include a 32 B prologue and a 32 B epilogue
        carry = os.environ.get("VPM_TUNE_VPCARRY", "1") != "0" and mode == 0
        kprof = min(args.steps, 20)
        call_bytes = BYTES_PER_STEP * kprof if carry else (BYTES_PER_STEP * (kprof - 1) + 64)
        if mode == 1:   # frozen field (the reference as shipped): one deposit pass per call (x, w: 16 B), then passes that read and
            call_bytes = 32 * kprof + 16   # write x, v only (32 B: without diagnostics the weights are not needed)
        bytes_per_launch = call_bytes * n / max(pass_cnt, 1)
        avg = pass_ms / max(pass_cnt, 1)
        achieved = bytes_per_launch / (avg * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("vp")
        kernel = ("vp_pass_ring_kernel (fused kick+drift+deposit; one pass per step, the stagger is carried from call to call)"
                  if os.environ.get("VPM_TUNE_TMA", "5") != "0" else "vp_pass_kernel")
        out = {"value": self.ntotal / (per * 1e-3), "ms_per_step": per, "launches": launches, "clocks": clocks,
               "passes": args.steps + (0 if carry else 1),
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
                            "traffic": traffic, "traffic_source": "one ncu --set full capture (profiles/traffic.json), not re-measured per run",
                            "peak_source": self.peak_src, "kernel": kernel, "bytes_per_launch": bytes_per_launch,
                            "avg_launch_ms": avg, "launches_timed": pass_cnt,
                            "field_kernel_share": field_ms / max(pass_ms + field_ms, 1e-30),
                            "frac_of_8TBs_nominal": achieved / 8000.0}}
        self.vp_run_steps = run_steps
        return out

    def bench_vp_uniform(self):
        """context only (NOT the headline): uniform-weight fast path, 32 B per particle-step.  Every reference sampler
        produces equal weights (w = L/N), so the steppers can skip the w[] stream when the caller declares it; `value`
        always streams w[] (40 B), as the reference's layout does."""
        vpm, d, n = self.vpm, self.vp_d, self.n
        d.set_uniform_weight(L / self.ntotal)
        self.vp_run_steps(3)
        ku = min(self.args.steps, 50)
        ms, _, _ = self.timed(self.vp_run_steps, ku, sample_clocks=False)
        ums = ms / ku
        vpm.initialize_(d, vpm.BumpOnTail(kappa=KAPPA), offset=self.rank * n, ntotal=self.ntotal)   # back to per-particle weights
        return {"value": self.ntotal / (ums * 1e-3), "unit": "particle-steps/s", "ms_per_step": ums,
                "bytes_per_particle_step": 32, "GBps": 32 * n / (ums * 1e-3) / 1e9,
                "api": "vpm_particles_set_uniform_weight + vpm_vp_strang_steps"}

    def pcie_ceiling(self, zin, zout, nbytes):
        """all ranks copy nbytes host->device, then device->host, at the same time (the e2e step's own pattern)"""
        import ctypes as C
        vpm, lib, ctx = self.vpm, self.lib, self.ctx
        dev = C.c_void_p()
        vpm.check(lib.vpm_dev_alloc(ctx._h, nbytes // 8, C.byref(dev)))
        res = {}
        for name, fn in (("h2d", lambda: lib.vpm_memcpy_h2d(ctx._h, dev, zin, nbytes // 8)),
                         ("d2h", lambda: lib.vpm_memcpy_d2h(ctx._h, zout, dev, nbytes // 8))):
            vpm.check(fn())
            self.barrier()
            t0 = time.perf_counter()
            vpm.check(fn())
            self.barrier()
            res[name] = self.max_over_ranks(time.perf_counter() - t0)
        vpm.check(lib.vpm_dev_free(ctx._h, dev))
        return res

    def bench_e2e(self):
        """the host-array drop-in step (z = 2 x N host matrix in, out), PCIe copies inside the timing"""
        import ctypes as C
        args, vpm, lib, n, d, pot = self.args, self.vpm, self.lib, self.n, self.vp_d, self.vp_pot
        zin, zout = C.c_void_p(), C.c_void_p()
        vpm.check(lib.vpm_host_alloc(16 * n, C.byref(zin)))
        vpm.check(lib.vpm_host_alloc(16 * n, C.byref(zout)))
        vpm.check(lib.vpm_particles_download_aos(d._h, zin, 2))
        vpm.check(lib.vpm_vp_strang_step_host(pot._h, d._h, zin, zout, DT, CHI, 0))  # warm-up (allocates staging)
        self.barrier()
        t0 = time.perf_counter()
        cur, nxt = zin, zout
        for _ in range(args.e2e_steps):
            vpm.check(lib.vpm_vp_strang_step_host(pot._h, d._h, cur, nxt, DT, CHI, 0))
            cur, nxt = nxt, cur
        self.barrier()
        t_e2e = self.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": self.ntotal * args.e2e_steps / t_e2e, "unit": "particle-steps/s",
               "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n, "steps": args.e2e_steps,
               "api": "vpm_vp_strang_step_host (pinned 2 x N integrator state in/out per step; weights resident)"}
        # the box's own ceiling for this traffic pattern: every rank moves the same 16 n bytes each way, concurrently
        try:
            pc = self.pcie_ceiling(zin, zout, 16 * n)
            ceil_step = pc["h2d"] + pc["d2h"]
            e2e["pcie"] = {"h2d_GBps_aggregate": 16 * n * self.world / pc["h2d"] / 1e9, "d2h_GBps_aggregate": 16 * n * self.world / pc["d2h"] / 1e9,
                           "copy_only_ms_per_step": 1e3 * ceil_step, "ms_per_step": 1e3 * t_e2e / args.e2e_steps,
                           "frac_of_pcie_ceiling": ceil_step / (t_e2e / args.e2e_steps), "host_numa_cpus": self.numa,
                           "note": "ceiling = the same bytes copied H2D then D2H by all ranks at once from the same pinned buffers, nothing else running"}
        except Exception as e:  # the ceiling is context, never fatal
            e2e["pcie"] = {"error": str(e)}
        # for context: what a run!(integrator) user gets — initial state uploaded once from pinned memory, K steps
        # on the device with (W,K,M) read back for every step, final state downloaded once
        kk = min(args.steps, 100)
        diag = np.zeros((kk + 1, 3))
        self.barrier()
        t0 = time.perf_counter()
        vpm.check(lib.vpm_particles_upload_aos(d._h, cur, 2))
        vpm.check(lib.vpm_vp_strang_steps(pot._h, d._h, DT, CHI, kk, 0, 1, diag.ctypes.data))
        vpm.check(lib.vpm_particles_download_aos(d._h, nxt, 2))
        self.barrier()
        t_run = self.max_over_ranks(time.perf_counter() - t0)
        e2e["run_api"] = {"value": self.ntotal * kk / t_run, "unit": "particle-steps/s", "steps": kk,
                          "h2d_bytes_total": 16 * n, "d2h_bytes_total": 16 * n + 24 * (kk + 1),
                          "api": "upload_aos + vpm_vp_strang_steps(diag_mode=1) + download_aos (state resident between steps)"}
        lib.vpm_host_free(zin)
        lib.vpm_host_free(zout)
        return e2e

    # ---------------------------------------------------------------------------------------------------- LB / CLB
    def bench_lb(self, cons, steps, sustained=True):
        """BASELINE configs[2] / [3]: RK438 steps of the (conservative) Lenard-Bernstein model, DoubleMaxwellian +-2,
        41 knots, order 4, nu 1, dt 1e-2 (scripts/lenard_bernstein{,_conservative}.jl)"""
        vpm, lib, ctx, n = self.vpm, self.lib, self.ctx, self.n
        d = vpm.ParticleDistribution(1, 1, n, ctx)
        vpm.initialize_(d, vpm.DoubleMaxwellian(LB_DOMAIN, LB_SHIFT), offset=self.rank * n, ntotal=self.ntotal)
        sd = vpm.SplineDistribution(1, 1, LB_NKNOTS, LB_ORDER, LB_DOMAIN, "Dirichlet", ctx)

        def run_steps(k):
            vpm.check(lib.vpm_lb_rk438_steps_async(sd._h, d._h, LB_NU, LB_DT, int(k), int(cons)))

        # the first stepper call builds the velocity-sorted mirror of (v, w) (once per ensemble: the collision flow keeps it
        # sorted); its cost is reported separately, and so is the write-back of v into the caller's order (paid when v is read)
        first_ms, _, _ = self.timed(run_steps, 3, sample_clocks=False)
        second_ms, _, _ = self.timed(run_steps, 3, sample_clocks=False)
        self.barrier()
        ms, launches, clocks = self.timed(run_steps, steps)
        per = ms / steps
        kms, kcnt, msl, cntl = self.profile(run_steps, min(steps, 10))
        sorted_path = os.environ.get("VPM_TUNE_LBSORT", "1") != "0" and n >= (1 << 18)   # csrc/cabi.cu lb_sort_mode
        passes = {}
        tot_ms, tot_bytes = 0.0, 0.0
        for mode, name in LB_PASS_NAME.items():
            if cntl[mode] == 0:
                continue
            # the private-histogram path of CLB stores q2, q3 for its moments passes; the velocity-sorted path (default) has none
            b = LB_PASS_BYTES[mode] + (8 if (cons and mode in (1, 2) and not sorted_path) else 0)
            avg = float(msl[mode]) / int(cntl[mode])
            gbs = b * n / (avg * 1e-3) / 1e9
            passes[name] = {"bytes_per_particle": b, "avg_launch_ms": avg, "launches_timed": int(cntl[mode]),
                            "GBps": gbs, "frac": gbs / self.peak}
            if mode != 0:
                tot_ms += float(msl[mode])
                tot_bytes += b * n * int(cntl[mode])
        bytes_step = 136 + (48 if (cons and not sorted_path) else 0)
        achieved = tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
        out = {"value": self.ntotal / (per * 1e-3), "unit": "particle-steps/s", "ms_per_step": per, "steps": steps,
               "config": {"workload": ("clb" if cons else "lb") + "_rk438_double_maxwellian", "particles_per_gpu": n,
                          "n_basis": LB_NKNOTS, "order": LB_ORDER, "dt": LB_DT, "nu": LB_NU},
               "bytes_per_particle_step": bytes_step, "gpu_launches": launches, "clocks": clocks,
               "whole_step_GBps": bytes_step * n / (per * 1e-3) / 1e9,
               "whole_step_frac": bytes_step * n / (per * 1e-3) / 1e9 / self.peak,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
                            "kernel": ("lbs_pass_kernel (velocity-sorted RK438 stage passes" if sorted_path else
                                       "lb_pass_ring_kernel (RK438 stage passes" + (" and moments passes" if cons else "")) + "; byte-weighted mean over the pass kinds)",
                            "field_kernel_share": float(kms[3]) / max(float(kms[2] + kms[3]), 1e-30),
                            "field_kernel_avg_ms": float(kms[3]) / max(int(kcnt[3]), 1)},
               "passes": passes}
        if sorted_path:
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            vpm.check(lib.vpm_particles_ptrs_const(d._h, None, C.byref(C.c_void_p()), None))   # reads v: triggers the write-back
            e1.record(self.stream)
            self.torch.cuda.synchronize()
            out["sorted_mirror"] = {"build_ms_once": max(first_ms - second_ms, 0.0), "writeback_ms_on_read": e0.elapsed_time(e1),
                                    "note": "the RK438 passes run on a velocity-sorted mirror of (v, w): built once per ensemble (a 1-D collision "
                                            "flow preserves the order), not inside the timed region; v returns to the caller's order when it is read"}
        if sorted_path and self.world == 1:
            # opt-in variant (every sampler of the reference produces equal weights): the caller declares it and the passes skip
            # the w stream: 104 instead of 136 B per particle-step; the mirror is rebuilt once (the declaration rewrites w)
            d.set_uniform_weight(1.0 / self.ntotal)
            run_steps(2)
            ums, _, _ = self.timed(run_steps, steps, sample_clocks=False)
            out["uniform_weight_variant"] = {"ms_per_step": ums / steps, "value": self.ntotal / (ums / steps * 1e-3), "unit": "particle-steps/s",
                                             "bytes_per_particle_step": 104, "GBps": 104 * n / (ums / steps * 1e-3) / 1e9,
                                             "frac": 104 * n / (ums / steps * 1e-3) / 1e9 / self.peak}
            vpm.initialize_(d, vpm.DoubleMaxwellian(LB_DOMAIN, LB_SHIFT), offset=self.rank * n, ntotal=self.ntotal)
            run_steps(2)
        if sustained:
            secs = self.args.sustained_seconds
            self.args.sustained_seconds = min(secs, 3.5) if self.args.workload == "vp" else secs   # short inside the default line
            out["sustained"] = self.sustained(run_steps, per, bytes_step)
            self.args.sustained_seconds = secs
        del d, sd
        return out

    # ---------------------------------------------------------------------------------------------------- parity
    def parity_check(self):
        """Config 5 must be CORRECT, not only fast: a small VP + CLB run on the same ranks / communicator as the timed
        region; every rank's field must be bitwise identical, and the gathered particles must reproduce (1e-12, normwise)
        a single-rank run of the same global ensemble on rank 0 and the CPU oracle."""
        vpm, torch, dist, ctx = self.vpm, self.torch, self.dist, self.ctx
        world, rank = self.world, self.rank
        nper, vp_steps, lb_steps = 300_000, 5, 3   # >= 2^18 per rank: the CLB steps take the velocity-sorted passes, as the timed region does
        ntot = nper * world

        def run_pair(c, npart, offset):
            d = vpm.ParticleDistribution(1, 1, npart, c)
            vpm.initialize_(d, vpm.BumpOnTail(kappa=KAPPA), offset=offset, ntotal=ntot)
            pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), ORDER, NH), c)
            m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(vp_steps, DT), DT, field="selfconsistent")
            vpm.run_(m, diag_mode=2)
            x, v, _ = d.get()
            d2 = vpm.ParticleDistribution(1, 1, npart, c)
            vpm.initialize_(d2, vpm.DoubleMaxwellian(LB_DOMAIN, LB_SHIFT), offset=offset, ntotal=ntot)
            sd = vpm.SplineDistribution(1, 1, LB_NKNOTS, LB_ORDER, LB_DOMAIN, "Dirichlet", c)
            gi = vpm.GeometricIntegrator(vpm.ConservativeLenardBernstein(d2, vpm.CollisionEntropy(sd), nu=LB_NU), vpm.tspan_for(lb_steps, LB_DT), LB_DT)
            vpm.run_(gi)
            return {"x": x, "v": v, "phi": pot.coefficients, "diag": m.diagnostics, "vlb": d2.get("v"),
                    "coef": sd.coefficients, "dlb": gi.diagnostics[:, :2].copy()}

        mine = run_pair(ctx, nper, rank * nper)
        if world > 1:
            small = {k: mine[k] for k in ("phi", "diag", "coef", "dlb")}
            allsmall = [None] * world
            dist.all_gather_object(allsmall, small)
            bitwise = all(np.array_equal(allsmall[0][k], a[k]) for a in allsmall for k in small)
            gathered = {}
            for k in ("x", "v", "vlb"):
                t = torch.from_numpy(mine[k]).cuda()
                parts = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(parts, t)
                gathered[k] = torch.cat(parts).cpu().numpy()
        else:
            bitwise, gathered = None, {k: mine[k] for k in ("x", "v", "vlb")}
        res = None
        if rank == 0:
            nrm = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
            res = {"particles_total": ntot, "vp_steps": vp_steps, "clb_steps": lb_steps, "tol": PARITY_TOL,
                   "ranks_bitwise_equal": bitwise}
            if world > 1:   # the same global ensemble on ONE rank: a second context on this GPU without a communicator
                c1 = vpm.Context(self.local, self.stream.cuda_stream)
                one = run_pair(c1, ntot, 0)
                dsc = np.abs(one["diag"]).max(axis=0) + 1e-300
                lsc = np.array([np.abs(one["vlb"]).sum(), (one["vlb"] ** 2).sum()])
                res["vs_single_rank"] = {"max_rel_x": nrm(gathered["x"], one["x"]), "max_rel_v": nrm(gathered["v"], one["v"]),
                                         "max_rel_phi": nrm(mine["phi"], one["phi"]),
                                         "max_rel_WK_history": float((np.abs(mine["diag"] - one["diag"]) / dsc)[:, :2].max()),
                                         "max_rel_clb_v": nrm(gathered["vlb"], one["vlb"]),
                                         "max_rel_clb_coef": nrm(mine["coef"], one["coef"]),
                                         "max_rel_clb_moments": float((np.abs(mine["dlb"] - one["dlb"]) / lsc).max())}
                del one
                c1.close()
            from oracle import oracle as orc
            orc.set_threads(orc.max_threads())
            x0, v0, w0 = orc.sample_bump_on_tail(ntot, kappa=KAPPA)
            xo, vo, do, phio = orc.XSpace(0.0, L, ORDER, NH).strang_selfconsistent(x0, v0, w0, DT, vp_steps, chi=CHI)
            _, vl0, wl0 = orc.sample_maxwellian(ntot, xlo=LB_DOMAIN[0], xhi=LB_DOMAIN[1], shift=LB_SHIFT, doubled=True)
            vlo, dlo = orc.VSpace(LB_DOMAIN[0], LB_DOMAIN[1], LB_NKNOTS, LB_ORDER).rk438(vl0, wl0, LB_NU, LB_DT, lb_steps, conservative=True)
            orc.set_threads(1)
            dsc = np.abs(do).max(axis=0) + 1e-300
            lsc = np.array([np.abs(vlo).sum(), (vlo ** 2).sum()])
            xso = orc.XSpace(0.0, L, ORDER, NH)
            res["vs_oracle"] = {"max_rel_x": nrm(gathered["x"], xo), "max_rel_v": nrm(gathered["v"], vo),
                                "max_rel_phi": nrm(mine["phi"], xso.poisson_solve(xso.deposit(xo, w0))),
                                "max_rel_WK_history": float((np.abs(mine["diag"] - do) / dsc)[:, :2].max()),
                                "max_rel_clb_v": nrm(gathered["vlb"], vlo),
                                "max_rel_clb_moments": float((np.abs(mine["dlb"] - dlo) / lsc).max())}
            errs = [v for blk in ("vs_single_rank", "vs_oracle") if blk in res for v in res[blk].values()]
            res["max_rel_x"] = max(res[b]["max_rel_x"] for b in ("vs_single_rank", "vs_oracle") if b in res)
            res["max_rel_v"] = max(res[b]["max_rel_v"] for b in ("vs_single_rank", "vs_oracle") if b in res)
            res["max_rel_phi"] = max(res[b]["max_rel_phi"] for b in ("vs_single_rank", "vs_oracle") if b in res)
            res["max_err"] = max(errs)
            res["ok"] = bool(max(errs) <= PARITY_TOL and (bitwise is not False) and all(np.isfinite(errs)))
        self.barrier()
        return res

    def finish(self):
        if self.world > 1:
            if self.comm_used == "p2p":
                self.ctx.p2p_check()
                self.dist.barrier()
                self.ctx.p2p_detach()
            else:
                self.ctx.comm_destroy()
            self.dist.destroy_process_group()


def main():
    global NH, ORDER, LB_NKNOTS
    args = parse()
    NH, ORDER, LB_NKNOTS = args.n_basis, args.order, args.lb_nknots
    if args.impl == "reference":
        return run_reference(args)

    B = Bench(args)
    world, rank, n = B.world, B.rank, B.n
    extras = not args.no_extras
    cores = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        cores = orc.max_threads()
    line = {"metric": "particle-steps/sec", "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic"}
    cfg = workload_config(args, world)
    line["comm"] = B.comm_used

    if args.workload == "vp":
        vp = B.bench_vp()
        line["passes_in_timed_region"] = vp["passes"]
        line.update({"value": vp["value"], "ms_per_step": vp["ms_per_step"], "config": cfg, "roofline": vp["roofline"],
                     "gpu_launches": vp["launches"], "clocks": vp["clocks"]})
        if (extras or args.sustained) and args.load == "bump_on_tail" and args.field == "selfconsistent":
            line["sustained"] = B.sustained(B.vp_run_steps, vp["ms_per_step"], BYTES_PER_STEP)
            # the fused pass alone, timed while the clocks are still at their sustained level
            kms, kcnt, _, _ = B.profile(B.vp_run_steps, 20)
            hot = float(kms[0]) / max(int(kcnt[0]), 1)
            line["sustained"]["pass_avg_launch_ms_hot"] = hot
            line["sustained"]["pass_frac_hot"] = BYTES_PER_STEP * n / (hot * 1e-3) / 1e9 / B.peak
        line["uniform_weight_variant"] = B.bench_vp_uniform() if world == 1 else None
        line["e2e"] = None if args.no_e2e else B.bench_e2e()
        line["cpu_baseline"] = cpu_baseline_block("vp", int(args.cpu_sample), cores) if cores else None
        if extras and args.load == "bump_on_tail" and args.field == "selfconsistent":
            del B.vp_d, B.vp_pot
            wl = {}
            for name, cons in (("lb", False), ("clb", True)):
                wl[name] = B.bench_lb(cons, args.lb_steps)
                wl[name]["cpu_baseline"] = cpu_baseline_block(name, int(args.cpu_sample), cores) if cores else None
            line["workloads"] = wl
    else:
        cons = args.workload == "clb"
        r = B.bench_lb(cons, args.steps, sustained=extras or args.sustained)
        # velocity-sorted path: four passes per step and nothing else (the projection is carried from the warm-up call);
        # histogram path: a deposit-only pass per call, and four moments passes per step for CLB
        line["passes_in_timed_region"] = 4 * args.steps if "sorted_mirror" in r else 4 * args.steps * (2 if cons else 1) + 1
        line.update({"value": r["value"], "ms_per_step": r["ms_per_step"], "config": cfg, "roofline": r["roofline"],
                     "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "passes": r["passes"],
                     "whole_step_frac": r["whole_step_frac"], "bytes_per_particle_step": r["bytes_per_particle_step"],
                     "sorted_mirror": r.get("sorted_mirror"), "sustained": r.get("sustained"), "e2e": None,
                     "cpu_baseline": cpu_baseline_block(args.workload, int(args.cpu_sample), cores) if cores else None})
    parity = B.parity_check() if extras else None
    line["parity_check"] = parity
    if rank == 0:
        print(json.dumps(line), flush=True)
    B.finish()
    if rank == 0 and parity is not None and not parity["ok"]:
        print("bench.py: parity_check FAILED: " + json.dumps(parity), file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
