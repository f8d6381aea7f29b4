"""Velocity-sorted Lenard-Bernstein / conservative Lenard-Bernstein RK438 passes (csrc/kernels_lbs.cu) against the oracle.

The collision flow cannot reorder particles in v, so the stepper sorts a mirror of (v, w) once and then deposits per-cell
power sums from registers; the field kernel turns them into the right-hand side and the five CLB moments
(src/projections/distribution.jl:35-55, src/projections/density.jl:43-52, src/models/lenard_bernstein_conservative.jl:11-36).
VPM_TUNE_LBSORT: 0 = private-histogram passes, 1 = default (sorted from 2^18 particles), 2 = sorted at any size,
3 = sorted passes on an UNSORTED mirror (every trip takes the mixed-cell path).  Tolerance: the north star's 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def nrm(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


@pytest.fixture(scope="module")
def vpm():
    import vpm_b200
    vpm_b200.default_context()
    return vpm_b200


def ensemble(n, seed, edge=True):
    rng = np.random.default_rng(seed)
    v = np.r_[rng.standard_normal(n // 2) + 2.0, rng.standard_normal(n - n // 2) - 2.0]
    rng.shuffle(v)
    if edge and n >= 6:
        v[:6] = [-10.0, 10.0, -10.5, 11.0, 9.999, -9.999]     # domain ends, out-of-domain particles, end cells
    w = rng.uniform(0.5, 1.5, n) / max(n, 1)
    return v, w


def run_gpu(vpm, sd, v, w, nu, dt, ns, cons, uniform=None, entropy=False):
    n = v.size
    d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), v, w)
    if uniform is not None:
        d.set_uniform_weight(uniform)
    model = (vpm.ConservativeLenardBernstein if cons else vpm.LenardBernstein)(d, vpm.CollisionEntropy(sd), nu=nu)
    gi = vpm.GeometricIntegrator(model, vpm.tspan_for(ns, dt), dt)
    vpm.run_(gi, entropy=entropy)
    return d, gi


@pytest.mark.parametrize("mode", ["2", "3"])
@pytest.mark.parametrize("n", [1, 2, 63, 511, 512, 513, 1537, 40_003])
def test_sorted_rk438_vs_oracle(vpm, oracle, perr, monkeypatch, n, mode):
    """ragged sizes around the 512-particle ring tile, weighted particles, both models; mode 3 feeds the sorted passes an
    unsorted mirror, so every trip goes through the per-distinct-cell shuffle reductions"""
    monkeypatch.setenv("VPM_TUNE_LBSORT", mode)
    v, w = ensemble(n, 300 + n)
    nu, dt, ns = 0.9, 0.02, 3
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    for cons in (False, True):
        if cons and n < 64:
            continue                                          # the CLB coefficients are 0/0 for a handful of particles
        d, gi = run_gpu(vpm, sd, v, w, nu, dt, ns, cons)
        vo, do = vs.rk438(v, w, nu, dt, ns, conservative=cons)
        mdl = "clb" if cons else "lb"
        perr(f"{mdl}_sorted_rk438_v@n{n}_m{mode}", np.abs(d.get("v") - vo).max() / max(1.0, np.abs(vo).max()), TOL)
        dscale = np.array([np.abs(v).sum(), (v * v).sum()])
        perr(f"{mdl}_sorted_rk438_moment_history@n{n}_m{mode}", (np.abs(gi.diagnostics[:, :2] - do) / dscale).max(), TOL)
        perr(f"{mdl}_sorted_spline_coefficients@n{n}_m{mode}", nrm(sd.coefficients, vs.project(vo, w)), TOL)


@pytest.mark.parametrize("cons", [False, True])
def test_sorted_uniform_weights_and_entropy(vpm, oracle, perr, monkeypatch, cons):
    """declared uniform weights (W = w S in the field kernel: no weight stream) and the entropy history on the sorted mirror"""
    monkeypatch.setenv("VPM_TUNE_LBSORT", "2")
    n, nu, dt, ns = 60_001, 0.9, 0.02, 4
    v, _ = ensemble(n, 7, edge=False)
    w = np.full(n, 1.0 / n)
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    vo, do, eo, nfo = vs.rk438_entropy(v, w, nu, dt, ns, conservative=cons, f_floor=vpm.ENTROPY_FLOOR)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    tag = "@clb" if cons else "@lb"
    for uniform in (None, 1.0 / n):
        d, gi = run_gpu(vpm, sd, v, w, nu, dt, ns, cons, uniform=uniform, entropy=True)
        u = "_uw" if uniform else ""
        perr("sorted_rk438_v" + u + tag, nrm(d.get("v"), vo), TOL)
        perr("sorted_entropy_history" + u + tag, np.abs(gi.entropy - eo).max() / np.abs(eo).max(), TOL)
        dscale = np.array([np.abs(v).sum(), (v * v).sum()])
        perr("sorted_moment_history" + u + tag, (np.abs(gi.diagnostics[:, :2] - do) / dscale).max(), TOL)


@pytest.mark.parametrize("K,nk", [(3, 33), (5, 12), (6, 25), (4, 101), (4, 201), (2, 64), (6, 140)])
def test_sorted_other_grids(vpm, oracle, perr, monkeypatch, K, nk):
    """other spline orders and knot counts (the power sums run up to u^(K+1)); tolerance: the mass matrix's backward-error bound"""
    monkeypatch.setenv("VPM_TUNE_LBSORT", "2")
    n, nu, dt, ns = 30_011, 0.7, 0.02, 2
    v, w = ensemble(n, 40 + K)
    vs = oracle.VSpace(-10.0, 10.0, nk, K)
    sd = vpm.SplineDistribution(1, 1, nk, K, (-10.0, 10.0), "Dirichlet")
    tol = max(TOL, 16.0 * np.linalg.cond(sd.mass_matrix) * 2.2e-16)
    for cons in (False, True):
        d, gi = run_gpu(vpm, sd, v, w, nu, dt, ns, cons)
        vo, do = vs.rk438(v, w, nu, dt, ns, conservative=cons)
        perr(f"{'clb' if cons else 'lb'}_sorted_rk438_v@K{K}_n{nk}", nrm(d.get("v"), vo), tol)


def test_sorted_matches_histogram_path_and_mirror_reuse(vpm, monkeypatch):
    """2e6 particles (default switch: sorted): sorted passes == private-histogram passes to rounding; a run split into two
    calls reuses the mirror and equals one run; anything that rewrites v in between invalidates the mirror"""
    n = 2_000_003
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    ent = vpm.CollisionEntropy(sd)
    d = vpm.ParticleDistribution(1, 1, n)
    vpm.initialize_(d, vpm.DoubleMaxwellian((-10.0, 10.0), 2.0))
    v0 = d.get("v")
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("VPM_TUNE_LBSORT", mode)
        for cons in (False, True):
            d.set(v=v0)
            model = (vpm.ConservativeLenardBernstein if cons else vpm.LenardBernstein)(d, ent, nu=1.0)
            gi = vpm.GeometricIntegrator(model, vpm.tspan_for(4, 1e-2), 1e-2)
            vpm.run_(gi)
            res[mode, cons] = (d.get("v"), gi.diagnostics.copy(), sd.coefficients.copy())
    for cons in (False, True):
        assert nrm(res["1", cons][0], res["0", cons][0]) < 1e-13
        assert nrm(res["1", cons][2], res["0", cons][2]) < 1e-12
        np.testing.assert_allclose(res["1", cons][1][:, 1], res["0", cons][1][:, 1], rtol=1e-13)
    monkeypatch.setenv("VPM_TUNE_LBSORT", "1")
    d.set(v=v0)
    for _ in range(2):                                        # 2 + 2 steps: the second call finds a valid mirror
        gi = vpm.GeometricIntegrator(vpm.ConservativeLenardBernstein(d, ent, nu=1.0), vpm.tspan_for(2, 1e-2), 1e-2)
        vpm.run_(gi)
    assert nrm(d.get("v"), res["1", True][0]) < 1e-14
    # a writer in between: the stale mirror must not be used
    d.set(v=v0)
    gi = vpm.GeometricIntegrator(vpm.ConservativeLenardBernstein(d, ent, nu=1.0), vpm.tspan_for(4, 1e-2), 1e-2)
    vpm.run_(gi)
    assert np.array_equal(d.get("v"), res["1", True][0])      # same order, same sums: bitwise reproducible
    dg = res["1", True][1]
    assert abs(dg[-1, 0] - dg[0, 0]) < 1e-9 * n and abs(dg[-1, 1] - dg[0, 1]) < 1e-9 * dg[0, 1]


def test_sorted_large_step_still_correct(vpm, oracle, perr, monkeypatch):
    """nu dt = 0.5: the stage maps are no longer guaranteed monotone, rows mix cells -- results must not depend on that"""
    monkeypatch.setenv("VPM_TUNE_LBSORT", "2")
    n, nu, dt, ns = 20_000, 5.0, 0.1, 2
    v, w = ensemble(n, 11)
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    for cons in (False, True):
        d, gi = run_gpu(vpm, sd, v, w, nu, dt, ns, cons)
        vo, do = vs.rk438(v, w, nu, dt, ns, conservative=cons)
        perr(f"{'clb' if cons else 'lb'}_sorted_rk438_v_large_step", nrm(d.get("v"), vo), 1e-11)


def test_sorted_trajectory_frames(vpm, oracle, perr, monkeypatch, tmp_path):
    """run! with trajectory output on the sorted path: frames are in the caller's particle order (write-back per leg)"""
    import h5mini
    monkeypatch.setenv("VPM_TUNE_LBSORT", "2")
    n, nu, dt, ns = 5_003, 0.9, 0.02, 4
    v, w = ensemble(n, 21)
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), v, w)
    gi = vpm.GeometricIntegrator(vpm.ConservativeLenardBernstein(d, vpm.CollisionEntropy(sd), nu=nu), vpm.tspan_for(ns, dt), dt)
    path = str(tmp_path / "clb.h5")
    vpm.run_(gi, path, save_stride=2)
    z = h5mini.File(path).read("z")
    assert z.shape == (3, n)
    np.testing.assert_array_equal(z[0], v)
    for f, k in ((1, 2), (2, 4)):
        vo, _ = vs.rk438(v, w, nu, dt, k, conservative=True)
        perr(f"clb_sorted_frame{f}", nrm(z[f], vo), TOL)


def test_sorted_mirror_follows_other_writers(vpm, oracle, perr, monkeypatch):
    """collisions, Vlasov-Poisson steps (the kick rewrites v: the mirror must be dropped and v synced first), collisions
    again -- against the same sequence on the oracle; then a caller holding WRITABLE device pointers (mirror rebuilt at every
    call from then on)"""
    monkeypatch.setenv("VPM_TUNE_LBSORT", "2")
    n, nu, dt = 30_001, 0.9, 0.02
    L = 2 * np.pi / 0.3
    rng = np.random.default_rng(5)
    x = rng.uniform(0.0, L, n)
    v, _ = ensemble(n, 31, edge=False)
    w = np.full(n, L / n)
    vs, xs = oracle.VSpace(-10.0, 10.0, 41, 4), oracle.XSpace(0.0, L, 4, 16)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    clb = vpm.ConservativeLenardBernstein(d, vpm.CollisionEntropy(sd), nu=nu)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), 4, 16))
    vo, xo = v, x
    for _ in range(2):
        vpm.run_(vpm.GeometricIntegrator(clb, vpm.tspan_for(2, dt), dt))
        vo, _ = vs.rk438(vo, w, nu, dt, 2, conservative=True)
        vpm.run_(vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(2, 0.1), 0.1, field="selfconsistent"))
        xo, vo, _, _ = xs.strang_selfconsistent(xo, vo, w, 0.1, 2)
    xg, vg, _ = d.get()
    perr("sorted_collisions_then_vp_x", nrm(xg, xo), TOL)
    perr("sorted_collisions_then_vp_v", nrm(vg, vo), TOL)
    # writable pointers handed out: the library can no longer know when v changes
    d.ptrs(writable=True)
    for _ in range(2):
        vpm.run_(vpm.GeometricIntegrator(clb, vpm.tspan_for(1, dt), dt))
        vo, _ = vs.rk438(vo, w, nu, dt, 1, conservative=True)
    perr("sorted_after_writable_pointers_v", nrm(d.get("v"), vo), TOL)


@pytest.mark.parametrize("case", ["one_cell", "identical", "all_outside"])
def test_sorted_degenerate_ensembles(vpm, oracle, perr, monkeypatch, case):
    """every particle in one cell (all CTAs feed the same columns), identical velocities (ties in the sort), nobody inside
    the spline domain (only the ghost row is touched: zero right-hand side)"""
    monkeypatch.setenv("VPM_TUNE_LBSORT", "2")
    n, nu, dt, ns = 20_000, 0.9, 0.02, 2
    rng = np.random.default_rng(9)
    v = {"one_cell": 0.26 + 0.2 * rng.random(n), "identical": np.full(n, 1.2345), "all_outside": 10.5 + rng.random(n)}[case]
    w = rng.uniform(0.5, 1.5, n) / n
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    d, gi = run_gpu(vpm, sd, v, w, nu, dt, ns, False)
    vo, do = vs.rk438(v, w, nu, dt, ns, conservative=False)
    perr("lb_sorted_rk438_v@" + case, np.abs(d.get("v") - vo).max() / np.abs(vo).max(), TOL)
    perr("lb_sorted_moment_history@" + case, (np.abs(gi.diagnostics[:, :2] - do) / np.array([np.abs(v).sum(), (v * v).sum()])).max(), TOL)


@pytest.mark.parametrize("cons", [False, True])
def test_sorted_projection_is_carried_between_calls(vpm, oracle, perr, monkeypatch, cons):
    """the stage-4 pass of a call leaves the projection of the final state solved: the next call on the same ensemble and
    space skips the deposit-only pass (4 passes + 4 field kernels per step, nothing else) and its history continues the last
    row; another ensemble using the same space in between voids that"""
    monkeypatch.setenv("VPM_TUNE_LBSORT", "2")
    n, nu, dt = 20_011, 0.9, 0.02
    v, w = ensemble(n, 77)
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), v, w)
    model = (vpm.ConservativeLenardBernstein if cons else vpm.LenardBernstein)(d, vpm.CollisionEntropy(sd), nu=nu)
    ctx = vpm.default_context()
    rows, counts = [], []
    for k in (1, 1, 2):
        l0 = ctx.launches
        gi = vpm.GeometricIntegrator(model, vpm.tspan_for(k, dt), dt)
        vpm.run_(gi)
        counts.append(ctx.launches - l0)
        rows.append(gi.diagnostics[:, :2] if not rows else gi.diagnostics[1:, :2])
        if len(rows) > 1:
            np.testing.assert_array_equal(gi.diagnostics[0, :2], last)      # row 0 of a carried call = the previous call's last row
        last = gi.diagnostics[-1, :2].copy()
    assert counts[1] == 8 and counts[2] == 16 and counts[0] > 8 + 2            # (first call: sort + deposit-only pass + its field kernel)
    vo, do = vs.rk438(v, w, nu, dt, 4, conservative=cons)
    tag = "@clb" if cons else "@lb"
    perr("sorted_carried_rk438_v" + tag, nrm(d.get("v"), vo), TOL)
    perr("sorted_carried_moment_history" + tag, (np.abs(np.concatenate(rows) - do) / np.array([np.abs(v).sum(), (v * v).sum()])).max(), TOL)
    # another ensemble projects into the same SplineDistribution: the carried projection is void
    other = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), v[::-1] * 0.5, w)
    vpm.projection(None, other, sd)
    l0 = ctx.launches
    vpm.run_(vpm.GeometricIntegrator(model, vpm.tspan_for(1, dt), dt))
    assert ctx.launches - l0 == 8 + 2
    vo, _ = vs.rk438(vo, w, nu, dt, 1, conservative=cons)
    perr("sorted_carry_voided_rk438_v" + tag, nrm(d.get("v"), vo), TOL)
