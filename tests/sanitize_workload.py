#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every particle
pass variant of the library at sizes that exercise whole ring tiles plus a ragged remainder, both histogram
fallbacks, the tiled large-grid deposit, the operators and the samplers.  Checks results against the oracle so
that a sanitizer-clean run is also a correct one.  Lives under tests/ because it uses the oracle as its checker
(only tests/, smoke() and bench.py's CPU arms may); it is a script, not a pytest module.

    compute-sanitizer --tool racecheck python tests/sanitize_workload.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_h5(vpm, orc, nrm):
    """run! drivers with trajectory output: snapshot passes, copy stream, pinned ring (several pieces), legs of steps"""
    import tempfile
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import h5mini
    n = (9 << 20) // 2 + 77      # frame = 2 n doubles > two 32 MiB ring pieces, with a ragged tail
    bot = vpm.BumpOnTail()
    x, v, w = orc.sample_bump_on_tail(n)
    with tempfile.TemporaryDirectory() as tmp:
        for field in ("selfconsistent", "frozen"):
            d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
            pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), 4, 16))
            m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), (0.0, 0.3), 0.1, field=field)
            path = os.path.join(tmp, field + ".h5")
            vpm.run_(m, path, save_stride=2, diag_mode=1)
            z = h5mini.File(path).read("z")
            xs = orc.XSpace(0.0, bot.L, 4, 16)
            if field == "selfconsistent":
                xo, vo, _, _ = xs.strang_selfconsistent(x, v, w, 0.1, 2)
            else:
                xo, vo, _ = xs.strang_frozen(x, v, x, w, 0.1, 2)
            assert z.shape == (3, n, 2) and nrm(z[1, :, 0], xo) < 1e-12 and nrm(z[1, :, 1], vo) < 1e-12, field
            xg, vg, _ = d.get()
            assert np.array_equal(z[2, :, 0], xg) and np.array_equal(z[2, :, 1], vg), field
        nl = 3 * 512 + 77
        vv = np.random.default_rng(1).standard_normal(nl) * 1.2
        ww = np.full(nl, 1.0 / nl)
        sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
        for cons in (False, True):
            d = vpm.ParticleDistribution(1, 1, nl).set(np.zeros(nl), vv, ww)
            model = (vpm.ConservativeLenardBernstein if cons else vpm.LenardBernstein)(d, vpm.CollisionEntropy(sd), nu=0.8)
            gi = vpm.GeometricIntegrator(model, (0.0, 0.06), 0.02)
            path = os.path.join(tmp, f"lb{int(cons)}.h5")
            vpm.run_(gi, path, save_stride=2)
            z = h5mini.File(path).read("z")
            vo, _ = orc.VSpace(-10.0, 10.0, 41, 4).rk438(vv, ww, 0.8, 0.02, 2, conservative=cons)
            assert z.shape == (3, nl) and np.abs(z[1] - vo).max() < 1e-10 and np.array_equal(z[2], d.get("v")), cons
    print("sanitize workload (run! with HDF5 output): all results match the oracle")


def main():
    import vpm_b200 as vpm
    from oracle import oracle as orc
    nrm = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    if "--run-h5" in sys.argv:
        return run_h5(vpm, orc, nrm)
    n = 3 * 1024 + 77
    # ---- Vlasov-Poisson: TMA main pass, prologue/epilogue passes, field kernel, both modes
    bot = vpm.BumpOnTail()
    x, v, w = orc.sample_bump_on_tail(n)
    for field in ("selfconsistent", "frozen"):
        d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
        pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), 4, 16))
        m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), (0.0, 0.3), 0.1, field=field)
        vpm.run_(m, diag_mode=1)
        xs = orc.XSpace(0.0, bot.L, 4, 16)
        if field == "selfconsistent":
            xo, vo, _, _ = xs.strang_selfconsistent(x, v, w, 0.1, 3)
        else:
            xo, vo, _ = xs.strang_frozen(x, v, x, w, 0.1, 3)
        xg, vg, _ = d.get()
        assert nrm(xg, xo) < 1e-12 and nrm(vg, vo) < 1e-12, field
    # the production stepper: no diagnostics, stagger carried across calls (edge passes with run-time pre-drift / mid-x store),
    # on the ring kernel (nh 16), the general-grid ring kernel (nh 17) and the tiled large-grid kernel (nh 400)
    for nh in (16, 17, 400):
        d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
        pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), 4, nh))
        for k in (1, 3, 1):
            vpm.run_(vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(k, 0.1), 0.1, field="selfconsistent"), diag_mode=0)
        xo, vo, _, _ = orc.XSpace(0.0, bot.L, 4, nh).strang_selfconsistent(x, v, w, 0.1, 5)
        xg, vg, _ = d.get()
        assert nrm(xg, xo) < 1e-12 and nrm(vg, vo) < 1e-12, ("carried stagger", nh)
    # large grid: tiled segmented-reduction deposit
    d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), 4, 400))
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), (0.0, 0.2), 0.1, field="selfconsistent")
    vpm.run_(m, diag_mode=1)
    xo, vo, _, _ = orc.XSpace(0.0, bot.L, 4, 400).strang_selfconsistent(x, v, w, 0.1, 2)
    assert nrm(d.get("x"), xo) < 1e-12
    # ---- Lenard-Bernstein: ring passes, register passes, histogram fallbacks, uniform weights
    vv = np.r_[np.random.default_rng(1).standard_normal(n - 4) * 1.2, [-10.0, 10.0, -10.5, 11.0]]
    ww = np.full(n, 1.0 / n)
    vs = orc.VSpace(-10.0, 10.0, 41, 4)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    # (default ring, 512-worker "fat" ring forced at this small size, register passes, ring for every mode, per-warp /
    # per-CTA CAS fallbacks, tile-sorted deposit)
    # and the velocity-sorted passes (kernels_lbs.cu): sorted mirror (2) and unsorted mirror (3: every trip on the mixed-cell path)
    for env in ({"VPM_TUNE_LBTMA": "-1"}, {"VPM_TUNE_LBFAT": "2"}, {"VPM_TUNE_LBTMA": "0"}, {"VPM_TUNE_LBTMA": "1022", "VPM_TUNE_LBFAT": "0"},
                {"VPM_TUNE_HM": "1"}, {"VPM_TUNE_HM": "2"}, {"VPM_TUNE_HM": "3"}, {"VPM_TUNE_LBSORT": "2"}, {"VPM_TUNE_LBSORT": "3"}):
        for k in ("VPM_TUNE_LBTMA", "VPM_TUNE_HM", "VPM_TUNE_LBFAT", "VPM_TUNE_LBSORT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for cons in (False, True):
            for uw in (False, True):
                d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), vv, ww)
                if uw:
                    d.set_uniform_weight(1.0 / n)
                model = (vpm.ConservativeLenardBernstein if cons else vpm.LenardBernstein)(d, vpm.CollisionEntropy(sd), nu=0.8)
                gi = vpm.GeometricIntegrator(model, (0.0, 0.04), 0.02)
                vpm.run_(gi)
                vo, _ = vs.rk438(vv, ww, 0.8, 0.02, 2, conservative=cons)
                assert np.abs(d.get("v") - vo).max() < 1e-10, (env, cons, uw)
    for k in ("VPM_TUNE_LBTMA", "VPM_TUNE_HM", "VPM_TUNE_LBFAT", "VPM_TUNE_LBSORT"):
        os.environ.pop(k, None)
    # sorted passes whose CTAs deposit nothing at all (every particle outside the spline domain: only the ghost row is touched)
    os.environ["VPM_TUNE_LBSORT"] = "2"
    vout = 10.5 + np.random.default_rng(2).random(n)
    d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), vout, ww)
    gi = vpm.GeometricIntegrator(vpm.LenardBernstein(d, vpm.CollisionEntropy(sd), nu=0.8), vpm.tspan_for(2, 0.02), 0.02)
    vpm.run_(gi)
    assert np.array_equal(d.get("v"), vout) and np.all(sd.coefficients == 0.0)   # f = f' = 0 outside the knots: nothing moves
    os.environ.pop("VPM_TUNE_LBSORT", None)
    # entropy history (gather pass with the replicated table + the ENT phase of the field kernel)
    d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), vv, ww)
    gi = vpm.GeometricIntegrator(vpm.ConservativeLenardBernstein(d, vpm.CollisionEntropy(sd), nu=0.8), vpm.tspan_for(2, 0.02), 0.02)
    vpm.run_(gi, entropy=True)
    _, _, eo, _ = vs.rk438_entropy(vv, ww, 0.8, 0.02, 2, conservative=True, f_floor=vpm.ENTROPY_FLOOR)
    assert np.abs(gi.entropy - eo).max() < 1e-11 * np.abs(eo).max()
    # the round-1 VP pass (CTA barrier per tile) next to the default warp-specialised ring, general and power-of-two grids
    for env, nh in (({"VPM_TUNE_TMA": "1"}, 16), ({"VPM_TUNE_TMA": "5", "VPM_TUNE_VPREL": "0"}, 16), ({}, 17)):
        os.environ.update(env)
        d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
        pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), 4, nh))
        m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(3, 0.1), 0.1, field="selfconsistent")
        vpm.run_(m, diag_mode=1)
        xo, vo, _, _ = orc.XSpace(0.0, bot.L, 4, nh).strang_selfconsistent(x, v, w, 0.1, 3)
        assert nrm(d.get("x"), xo) < 1e-12 and nrm(d.get("v"), vo) < 1e-12, (env, nh)
        for k in env:
            os.environ.pop(k, None)
    # operators and samplers
    d = vpm.ParticleDistribution(1, 1, n)
    vpm.initialize_(d, vpm.DoubleMaxwellian((-10.0, 10.0), 2.0))
    fs = vpm.projection(None, d, sd)
    fs(d.get("v"))
    vpm.compute_f_densities(sd, d.get("v"))
    vpm.initialize_(d, vpm.NormalDistribution((0.0, 1.0)))
    vpm.initialize_(d, vpm.UniformDistribution())
    print("sanitize workload: all results match the oracle")


if __name__ == "__main__":
    main()
