"""run!(method, h5file) on the device: vpm_vp_run / vpm_lb_run step the particles, snapshot every
save_stride-th state device-to-device and stream it to an HDF5 file in the reference's layout
(src/methods/splitting.jl:23-52, src/methods/geometric_integrator.jl:12-44) while the next steps compute.
Checked against the CPU oracle at the saved steps and against the same run without output."""
import os

import numpy as np
import pytest

import h5mini

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


def _vp_setup(vpm, oracle, n, K=4, nh=16):
    bot = vpm.BumpOnTail()
    x, v, w = oracle.sample_bump_on_tail(n)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), K, nh))
    return bot, x, v, w, pot


@pytest.mark.parametrize("stride", [1, 3])
def test_vp_selfconsistent_trajectory(vpm, oracle, tmp_path, stride):
    n, nt, dt = 20011, 7, 0.1
    bot, x, v, w, pot = _vp_setup(vpm, oracle, n)
    d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(nt, dt), dt, field="selfconsistent")
    path = tmp_path / "vp.h5"
    vpm.run_(m, str(path), save_stride=stride, diag_mode=1)
    steps = sorted(set(list(range(0, nt, stride)) + [nt]))
    assert m.frames == len(steps)
    f = h5mini.File(path)
    z = f.read("z")
    assert z.shape == (len(steps), n, 2) and f.datasets["z"].chunk == (1, n, 2) and f.datasets["z"].maxshape == (None, n, 2)
    np.testing.assert_allclose(f.read("t"), dt * np.array(steps), rtol=1e-15)
    np.testing.assert_array_equal(z[0], np.stack([x, v], axis=1))       # frame 0 = initial conditions, bit for bit
    xs = oracle.XSpace(0.0, bot.L, 4, 16)
    for fr, s in enumerate(steps):
        xo, vo, _, _ = xs.strang_selfconsistent(x, v, w, dt, s)
        assert nrm(z[fr, :, 0], xo) < TOL and nrm(z[fr, :, 1], vo) < TOL, (fr, s)
    xg, vg, _ = d.get()
    np.testing.assert_array_equal(z[-1], np.stack([xg, vg], axis=1))    # the last frame is the final device state
    # the same run without output: legs of `stride` steps must not change the arithmetic or the history
    d2 = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    m2 = vpm.SplittingMethod(vpm.VlasovPoisson(d2, pot), vpm.tspan_for(nt, dt), dt, field="selfconsistent")
    vpm.run_(m2, diag_mode=1)
    x2, v2, _ = d2.get()
    assert nrm(xg, x2) < 1e-14 and nrm(vg, v2) < 1e-14
    assert m.diagnostics.shape == (nt + 1, 3)
    np.testing.assert_allclose(m.diagnostics, m2.diagnostics, rtol=1e-12)


def test_vp_frozen_field_stays_frozen_across_legs(vpm, oracle, tmp_path):
    """run!(::SplittingMethod) deposits the field from model.distribution, which does not move during
    the run; saving frames (legs of save_stride steps) must not refresh it."""
    n, nt, dt = 10007, 6, 0.1
    bot, x, v, w, pot = _vp_setup(vpm, oracle, n, K=3)
    out = []
    for h5 in (None, str(tmp_path / "frozen.h5")):
        d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
        m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(nt, dt), dt, field="frozen")
        vpm.run_(m, h5, save_stride=2, diag_mode=1)
        out.append((d.get(), m.diagnostics))
    (xa, va, _), da = out[0]
    (xb, vb, _), db = out[1]
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(va, vb)
    np.testing.assert_allclose(da, db, rtol=1e-13)
    z = h5mini.File(tmp_path / "frozen.h5").read("z")
    assert z.shape == (4, n, 2)
    xs = oracle.XSpace(0.0, bot.L, 3, 16)
    xo, vo = xs.strang_frozen(x, v, x, w, dt, 4)[:2]
    assert nrm(z[2, :, 0], xo) < TOL and nrm(z[2, :, 1], vo) < TOL


@pytest.mark.parametrize("cons", [False, True])
def test_lb_trajectory(vpm, oracle, tmp_path, cons):
    n, nt, dt, nu, t0 = 30001, 5, 0.02, 0.9, 1.5
    rng = np.random.default_rng(7)
    v = rng.standard_normal(n) * 1.2 + 0.3
    w = np.full(n, 1.0 / n)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), v, w)
    model = (vpm.ConservativeLenardBernstein if cons else vpm.LenardBernstein)(d, vpm.CollisionEntropy(sd), nu=nu)
    gi = vpm.GeometricIntegrator(model, vpm.tspan_for(nt, dt, t0), dt)
    path = tmp_path / "lb.h5"
    vpm.run_(gi, str(path), save_stride=2)
    f = h5mini.File(path)
    z, t = f.read("z"), f.read("t")
    steps = [0, 2, 4, 5]
    assert gi.frames == 4 and z.shape == (4, n) and f.datasets["z"].chunk == (1, n) and f.datasets["t"].chunk == (1,)
    np.testing.assert_allclose(t, t0 + dt * np.array(steps), rtol=1e-15)
    np.testing.assert_array_equal(z[0], v)
    np.testing.assert_array_equal(z[-1], d.get("v"))
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    vo, do = vs.rk438(v, w, nu, dt, nt, conservative=cons)
    for fr, s in enumerate(steps[1:], start=1):
        vs_, _ = vs.rk438(v, w, nu, dt, s, conservative=cons)
        assert np.abs(z[fr] - vs_).max() <= 1e-11 * np.abs(vs_).max(), (fr, s)
    np.testing.assert_allclose(gi.diagnostics, do, rtol=1e-10, atol=1e-12)


def test_large_frames_stream_through_the_pinned_ring(vpm, tmp_path):
    """1.2e7 particles: a frame is 192 MB = six 32 MiB pieces through four pinned slots, on the copy stream, while
    the next step runs; the file must hold exactly the device state at every step."""
    n, dt = 12_000_001, 0.1
    bot = vpm.BumpOnTail()
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), 4, 16))
    d = vpm.initialize_(vpm.ParticleDistribution(1, 1, n), bot)
    x0, v0, _ = d.get()
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(2, dt), dt, field="selfconsistent")
    path = tmp_path / "big.h5"
    vpm.run_(m, str(path), save_stride=1, diag_mode=0)
    assert os.path.getsize(path) > 3 * 16 * n
    z = h5mini.File(path).read("z")
    np.testing.assert_array_equal(z[0, :, 0], x0)
    np.testing.assert_array_equal(z[0, :, 1], v0)
    x2, v2, _ = d.get()
    np.testing.assert_array_equal(z[2, :, 0], x2)
    np.testing.assert_array_equal(z[2, :, 1], v2)
    # frame 1 = one step from the initial state
    d1 = vpm.initialize_(vpm.ParticleDistribution(1, 1, n), bot)
    m1 = vpm.SplittingMethod(vpm.VlasovPoisson(d1, pot), vpm.tspan_for(1, dt), dt, field="selfconsistent")
    vpm.run_(m1, diag_mode=0)
    x1, v1, _ = d1.get()
    np.testing.assert_array_equal(z[1, :, 0], x1)
    np.testing.assert_array_equal(z[1, :, 1], v1)
