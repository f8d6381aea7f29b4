"""Uniform-weight fast path (vpm_particles_set_uniform_weight): the steppers skip the w[] stream; results must be
bit-identical to the general path with the same (constant) weights, and match the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def nrm(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def vpm():
    import vpm_b200
    vpm_b200.default_context()
    return vpm_b200


@pytest.mark.parametrize("n", [20001, 4096])
def test_vp_uniform_weight(vpm, oracle, n):
    L = 2 * np.pi / 0.3
    x, v, w = oracle.sample_bump_on_tail(n)
    assert np.all(w == w[0])
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), 4, 16))
    out = []
    for uw in (False, True):
        d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
        if uw:
            d.set_uniform_weight(w[0])
        for field in ("selfconsistent", "frozen"):
            m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(3, 0.1), 0.1, field=field)
            vpm.run_(m, diag_mode=1)
        out.append(d.get() + (m.diagnostics,))
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)
    xs = oracle.XSpace(0.0, L, 4, 16)
    xo, vo, _, _ = xs.strang_selfconsistent(x, v, w, 0.1, 3)
    xo, vo, _ = xs.strang_frozen(xo, vo, xo, w, 0.1, 3)
    assert nrm(out[1][0], xo) < 1e-12 and nrm(out[1][1], vo) < 1e-12
    # uploading weights clears the declaration
    d.set(w=np.linspace(0.5, 1.5, n) * w[0])
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(1, 0.1), 0.1, field="selfconsistent")
    vpm.run_(m, diag_mode=2)
    x2, v2, w2 = d.get()
    xo2, vo2, do2, _ = xs.strang_selfconsistent(out[1][0], out[1][1], w2, 0.1, 1)
    assert nrm(x2, xo2) < 1e-12 and nrm(v2, vo2) < 1e-12


def test_lb_uniform_weight(vpm, oracle):
    n = 30001
    rng = np.random.default_rng(9)
    v = np.r_[rng.standard_normal(n // 2) + 2, rng.standard_normal(n - n // 2) - 2]
    w = np.full(n, 1.0 / n)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    res = []
    for uw in (False, True):
        d = vpm.ParticleDistribution(1, 1, n).set(np.zeros(n), v, w)
        if uw:
            d.set_uniform_weight(1.0 / n)
        gi = vpm.GeometricIntegrator(vpm.ConservativeLenardBernstein(d, vpm.CollisionEntropy(sd)), vpm.tspan_for(3, 0.01), 0.01)
        vpm.run_(gi)
        res.append((d.get("v"), gi.diagnostics))
    # the w[] stream changes which pass variants fit (ring vs register prefetch), hence the summation order
    assert nrm(res[0][0], res[1][0]) < 1e-13
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-12)
    vo, do = oracle.VSpace(-10.0, 10.0, 41, 4).rk438(v, w, 1.0, 0.01, 3, conservative=True)
    assert nrm(res[1][0], vo) < 1e-11


def test_writable_weight_pointer_ends_the_declaration(vpm, oracle):
    """Handing out the WRITABLE w pointer (vpm_particles_ptrs) ends a uniform-weight declaration -- a caller who rewrites
    the weights through it must not keep getting the old uniform weight from the steppers; the read-only accessor that the
    operators use (vpm_particles_ptrs_const) leaves it alone."""
    n = 20001
    bot = vpm.BumpOnTail()
    x, v, w = oracle.sample_bump_on_tail(n)
    d = vpm.ParticleDistribution(1, 1, n).set(x, v, w)
    d.set_uniform_weight(w[0])
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), 4, 16))
    vpm.projection_(pot, d)                                   # read-only pointers: declaration still active
    xp, vp, wp = d.ptrs(writable=True)                        # the caller may now rewrite w ...
    wnew = np.linspace(0.5, 1.5, n) * w[0]
    lib = vpm._cabi.lib()
    vpm.check(lib.vpm_memcpy_h2d(d.ctx._h, wp, wnew.ctypes.data, n))   # ... and does, behind the library's back
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(2, 0.1), 0.1, field="selfconsistent")
    vpm.run_(m, diag_mode=2)
    xs = oracle.XSpace(0.0, bot.L, 4, 16)
    xo, vo, do, _ = xs.strang_selfconsistent(x, v, wnew, 0.1, 2)
    xg, vg, _ = d.get()
    assert nrm(xg, xo) < 1e-12 and nrm(vg, vo) < 1e-12
