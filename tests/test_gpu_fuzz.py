"""Randomised parity sweep: random orders, grid sizes, domains, particle counts (odd, tiny, non-multiples of
the vector width), time steps and chi against the oracle.  Seeds are fixed, so failures are reproducible."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12
EPS = 2.220446049250313e-16


def nrm(a, b):
    d = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (d if d > 0 else 1.0)


@pytest.fixture(scope="module")
def vpm():
    import vpm_b200
    vpm_b200.default_context()
    return vpm_b200


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_vlasov_poisson(vpm, oracle, perr, seed):
    rng = np.random.default_rng(1000 + seed)
    K = int(rng.integers(2, 7))
    nh = int(rng.choice([3, 5, 8, 16, 17, 31, 64, 129, 300]))
    lo = float(rng.uniform(-5, 5))
    L = float(rng.uniform(0.5, 30.0))
    n = int(rng.choice([1, 2, 3, 31, 255, 257, 1000, 4097, 20001]))
    dt = float(rng.uniform(0.01, 0.3))
    chi = float(rng.choice([1.0, 0.7, 2.5]))
    ns = int(rng.integers(1, 5))
    x = rng.uniform(lo - 3 * L, lo + 4 * L, n)
    v = rng.standard_normal(n) * rng.uniform(0.1, 3.0)
    w = rng.uniform(0.1, 2.0, n) / n
    xs = oracle.XSpace(lo, lo + L, K, nh)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((lo, lo + L), K, nh))
    for field, dm in (("selfconsistent", 2), ("selfconsistent", 1), ("frozen", 0)):
        d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
        m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(ns, dt), dt, field=field, chi=chi if field != "frozen" else 1.0)
        vpm.run_(m, diag_mode=dm)
        xg, vg, _ = d.get()
        if field == "frozen":
            xo, vo, _ = xs.strang_frozen(x, v, x, w, dt, ns)
        else:
            xo, vo, do, _ = xs.strang_selfconsistent(x, v, w, dt, ns, chi=chi)
        tag = f"@seed{seed}_K{K}_nh{nh}_n{n}_{field}_dm{dm}"
        perr("fuzz_strang_x" + tag, nrm(xg, xo), TOL)
        perr("fuzz_strang_v" + tag, np.linalg.norm(vg - vo) / max(np.linalg.norm(vo), 1e-3 * np.sqrt(n)), TOL)
        if dm == 2:
            # W = phi' S phi / 2 of the solved field: backward-error bound of the solve on this (random) grid;
            # K relative to itself; M = sum w v relative to sum w |v|
            M_, S_ = xs.matrices()
            ev = np.linalg.eigvalsh(S_)
            kS = ev[-1] / ev[1] if nh > 1 else 1.0
            scale = np.array([np.abs(do[:, 0]).max(), np.abs(do[:, 1]).max(), np.abs(w * v).sum()]) + 1e-300
            e = (np.abs(m.diagnostics - do) / scale).max(axis=0)
            perr("fuzz_W_history" + tag, e[0], max(TOL, 64 * kS * EPS))
            perr("fuzz_K_history" + tag, e[1], TOL)
            perr("fuzz_M_history" + tag, e[2], TOL)


@pytest.mark.parametrize("ring", ["-1", "0"])   # TMA ring passes (default) / register-prefetch passes
@pytest.mark.parametrize("seed", range(8))
def test_fuzz_lenard_bernstein(vpm, oracle, perr, seed, ring, monkeypatch):
    monkeypatch.setenv("VPM_TUNE_LBTMA", ring)
    rng = np.random.default_rng(2000 + seed)
    K = int(rng.integers(3, 7))
    nk = int(rng.choice([8, 12, 21, 41, 64, 150]))
    lo, hi = -float(rng.uniform(6, 12)), float(rng.uniform(6, 12))
    n = int(rng.choice([2, 33, 1001, 4096, 30001]))
    nu = float(rng.uniform(0.2, 2.0))
    dt = float(rng.uniform(1e-3, 5e-2))
    ns = int(rng.integers(1, 4))
    cons = bool(rng.integers(0, 2)) and n > 100
    v = np.r_[rng.standard_normal(n // 2) * 0.8 + 1.5, rng.standard_normal(n - n // 2) * 1.2 - 1.0]
    if n > 30:
        v[:4] = [lo, hi, lo - 1.0, hi + 2.0]          # domain ends and out-of-domain particles
    w = rng.uniform(0.5, 1.5, n) / n
    vs = oracle.VSpace(lo, hi, nk, K)
    sd = vpm.SplineDistribution(1, 1, nk, K, (lo, hi), "Dirichlet")
    d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), v, w)
    model = (vpm.ConservativeLenardBernstein if cons else vpm.LenardBernstein)(d, vpm.CollisionEntropy(sd), nu=nu)
    gi = vpm.GeometricIntegrator(model, vpm.tspan_for(ns, dt), dt)
    vpm.run_(gi)
    vo, do = vs.rk438(v, w, nu, dt, ns, conservative=cons)
    tag = f"@seed{seed}_K{K}_nk{nk}_n{n}_{'clb' if cons else 'lb'}_ring{ring}"
    tol = max(TOL, 64 * np.linalg.cond(vs.mass()) * EPS)     # random stress grids: backward-error bound of the mass solve
    perr("fuzz_rk438_v" + tag, nrm(d.get("v"), vo), tol)
    dscale = np.array([np.abs(v).sum(), (v * v).sum()])
    perr("fuzz_rk438_moment_history" + tag, (np.abs(gi.diagnostics[:, :2] - do) / dscale).max(), tol)
