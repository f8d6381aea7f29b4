"""Carried stagger of the self-consistent Strang stepper (csrc/cabi.cu, vp_steps_carry): without diagnostics the last pass of
a stepper call is a full fused pass (kick + drift/2 | drift/2 + deposit) that also stores the caller-visible position, so
the next call needs no prologue pass -- as long as nobody touched the particles or the field in between.  Steps of the
legacy loop: src/vlasov_poisson.jl:58-67 (restated in oracle/vpm_oracle.c).  Tolerance: the north star's 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12
KAPPA = 0.3
L = 2 * np.pi / KAPPA


def nrm(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def vpm():
    import vpm_b200
    vpm_b200.default_context()
    return vpm_b200


def stepper(vpm, d, pot, k, dt=0.1):
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(k, dt), dt, field="selfconsistent")
    vpm.run_(m, diag_mode=0)


@pytest.mark.parametrize("n", [1537, 200_003])
@pytest.mark.parametrize("nh", [16, 17])
def test_repeated_calls_match_oracle(vpm, oracle, perr, monkeypatch, n, nh):
    """6 steps as one call, as 1 + 2 + 3, and as six single-step calls (one pass per step once the stagger is carried),
    with and without the carry: all equal to the oracle, and to each other to rounding"""
    x, v, w = oracle.sample_bump_on_tail(n)
    xs = oracle.XSpace(0.0, L, 4, nh)
    xo, vo, _, _ = xs.strang_selfconsistent(x, v, w, 0.1, 6)
    res = {}
    for carry in ("1", "0"):
        monkeypatch.setenv("VPM_TUNE_VPCARRY", carry)
        for name, calls in (("one", [6]), ("split", [1, 2, 3]), ("single", [1] * 6)):
            d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
            pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), 4, nh))
            l0 = vpm.default_context().launches
            for k in calls:
                stepper(vpm, d, pot, k)
            res[carry, name] = d.get() + (vpm.default_context().launches - l0,)
            tag = f"@n{n}_nh{nh}_carry{carry}_{name}"
            perr("carry_x" + tag, nrm(res[carry, name][0], xo), TOL)
            perr("carry_v" + tag, nrm(res[carry, name][1], vo), TOL)
    # launches (passes + field kernels): carried: 1 prologue + 6 fused, whatever the split; two-ended: calls + 6
    assert res["1", "single"][3] == res["1", "one"][3] == 2 * 7
    assert res["0", "single"][3] == 2 * 12 - 6 and res["0", "one"][3] == 2 * 7 - 1   # (the epilogue pass has no field kernel)
    assert nrm(res["1", "single"][0], res["0", "one"][0]) < 1e-13


def test_stagger_is_dropped_when_something_else_touches_state(vpm, oracle, perr, monkeypatch):
    """writers of x, v, w, another ensemble depositing into the same Potential, a diagnostics call, a uniform-weight
    declaration: each must void the carried stagger (the results say so: a stale stagger is an O(1) error)"""
    monkeypatch.setenv("VPM_TUNE_VPCARRY", "1")
    n = 50_001
    x, v, w = oracle.sample_bump_on_tail(n)
    xs = oracle.XSpace(0.0, L, 4, 16)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), 4, 16))
    d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    other = vpm.ParticleDistribution(1, 1, n).set((x + 1.0) % L, v[::-1].copy(), w)
    xo, vo = x, v
    rng = np.random.default_rng(3)

    def both(k):
        nonlocal xo, vo
        stepper(vpm, d, pot, k)
        xo, vo, _, _ = xs.strang_selfconsistent(xo, vo, w, 0.1, k)

    both(2)
    both(1)                                                   # carried
    xo = (xo + 0.3 * rng.random(n)) % L
    d.set(x=xo)                                               # x rewritten
    both(2)
    vo = vo * 0.9
    d.set(v=vo)                                               # v rewritten
    both(1)
    stepper(vpm, other, pot, 1)                               # the field in pot now belongs to another ensemble
    both(2)
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(1, 0.1), 0.1, field="selfconsistent")
    vpm.run_(m, diag_mode=2)                                  # a diagnostics call (two-ended scheme) in between
    xo, vo, _, _ = xs.strang_selfconsistent(xo, vo, w, 0.1, 1)
    both(1)
    stepper(vpm, d, pot, 1, dt=0.05)                          # another time step: the stagger was taken with dt = 0.1
    xo, vo, _, _ = xs.strang_selfconsistent(xo, vo, w, 0.05, 1)
    both(1)
    xg, vg, _ = d.get()
    perr("carry_dropped_x", nrm(xg, xo), TOL)
    perr("carry_dropped_v", nrm(vg, vo), TOL)
    # uniform weights declared between two calls (same values here, but the deposit switches kernels)
    wu = np.full(n, L / n)
    d.set(w=wu)
    stepper(vpm, d, pot, 1)
    d.set_uniform_weight(L / n)
    stepper(vpm, d, pot, 2)
    stepper(vpm, d, pot, 1)
    xo, vo, _, _ = xs.strang_selfconsistent(xo, vo, wu, 0.1, 4)
    xg, vg, _ = d.get()
    perr("carry_uniform_weights_x", nrm(xg, xo), TOL)
    perr("carry_uniform_weights_v", nrm(vg, vo), TOL)


@pytest.mark.parametrize("n", [0, 1, 2, 511, 513])
def test_carry_tiny_ensembles(vpm, oracle, perr, monkeypatch, n):
    """empty and tiny inputs (no whole ring tile): the edge passes run on the run-time-flag kernels"""
    monkeypatch.setenv("VPM_TUNE_VPCARRY", "1")
    rng = np.random.default_rng(n)
    x, v, w = rng.uniform(0.0, L, n), rng.standard_normal(n), np.full(n, L / max(n, 1))
    d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), 4, 16))
    for k in (1, 2, 1):
        stepper(vpm, d, pot, k)
    if n == 0:
        return
    xo, vo, _, _ = oracle.XSpace(0.0, L, 4, 16).strang_selfconsistent(x, v, w, 0.1, 4)
    xg, vg, _ = d.get()
    perr(f"carry_tiny_x@n{n}", np.abs(xg - xo).max() / max(1.0, np.abs(xo).max()), TOL)
    perr(f"carry_tiny_v@n{n}", np.abs(vg - vo).max() / max(1.0, np.abs(vo).max()), TOL)


def test_run_with_trajectory_output_carries_the_stagger(vpm, oracle, perr, monkeypatch, tmp_path):
    """run!(method, h5file) without diagnostics: the legs between saved frames carry the stagger (one pass per step plus
    the snapshot pass of each frame); frames in the file are the caller-visible states"""
    import h5mini
    monkeypatch.setenv("VPM_TUNE_VPCARRY", "1")
    n = 30_001
    x, v, w = oracle.sample_bump_on_tail(n)
    d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), 4, 16))
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(5, 0.1), 0.1, field="selfconsistent")
    path = str(tmp_path / "vp.h5")
    l0 = vpm.default_context().launches
    vpm.run_(m, path, save_stride=2, diag_mode=0)
    launches = vpm.default_context().launches - l0
    z = h5mini.File(path).read("z")
    assert z.shape == (4, n, 2)
    xs = oracle.XSpace(0.0, L, 4, 16)
    for f, k in enumerate((0, 2, 4, 5)):
        xo, vo, _, _ = xs.strang_selfconsistent(x, v, w, 0.1, k)
        perr(f"carry_run_frame{f}_x", nrm(z[f, :, 0], xo), TOL)
        perr(f"carry_run_frame{f}_v", nrm(z[f, :, 1], vo), TOL)
    # 1 deposit-only pass + 5 fused passes, 6 field kernels, 4 snapshot passes
    assert launches == 1 + 5 + 6 + 4, launches
