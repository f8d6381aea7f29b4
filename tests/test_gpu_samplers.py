"""Device-side NormalDistribution sampler (src/examples/normal.jl:10-36) against its CPU twin, and the
config-1 script (scripts/vlasov_poisson.jl as shipped: 1e4 particles, 16 knots, order 3, frozen-field Strang)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vpm():
    import vpm_b200
    vpm_b200.default_context()
    return vpm_b200


def test_normal_distribution_sampler(vpm, oracle):
    n = 10000
    d = vpm.ParticleDistribution(1, 1, n)
    vpm.initialize_(d, vpm.NormalDistribution((0.0, 1.0)))
    xo, vo, wo, xmax = oracle.sample_normal(n)
    x, v, w = d.get()
    assert d.xmax == xmax and xmax == np.ceil(xmax) and 3 <= xmax <= 6
    np.testing.assert_allclose(x, xo, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(v, vo, rtol=1e-13, atol=1e-14)
    np.testing.assert_array_equal(w, wo)
    assert x.min() >= 0.0 and x.max() <= 1.0 and abs(x.mean() - 0.5) < 0.01
    # a slab of a larger ensemble with an agreed xmax (multi-rank usage)
    d2 = vpm.ParticleDistribution(1, 1, 1000)
    vpm.initialize_(d2, vpm.NormalDistribution((-2.0, 3.0)), offset=5000, ntotal=n, xmax=5.0)
    xo2, vo2, wo2, _ = oracle.sample_normal(1000, offset=5000, Ntotal=n, xlo=-2.0, xhi=3.0, xmax=5.0)
    np.testing.assert_allclose(d2.get("x"), xo2, rtol=1e-13, atol=1e-14)


def test_uniform_and_shifted_samplers(vpm, oracle):
    """UniformDistribution, ShiftedUniformDistribution, ShiftedNormalV (src/examples/{uniform,shifteduniform,
    shiftednormalv}.jl) against their CPU twins, as slabs of a larger ensemble."""
    n, off, ntot = 4001, 1234, 20000
    d = vpm.ParticleDistribution(1, 1, n)
    vpm.initialize_(d, vpm.UniformDistribution((0.5, 2.5), (-3.0, 1.0)), offset=off, ntotal=ntot)
    xo, vo, wo = oracle.sample_uniform(n, off, ntot, xlo=0.5, xhi=2.5, vlo=-3.0, vhi=1.0)
    x, v, w = d.get()
    np.testing.assert_allclose(x, xo, rtol=1e-14)
    np.testing.assert_allclose(v, vo, rtol=1e-14, atol=1e-15)
    np.testing.assert_array_equal(w, wo)
    assert 0.5 <= x.min() and x.max() < 2.5 and -3.0 <= v.min() and v.max() < 1.0 and w[0] == 1.0 / ntot
    vpm.initialize_(d, vpm.ShiftedUniformDistribution((0.0, 1.0), (-2.0, 2.0), 2.0), offset=off, ntotal=ntot)
    xo, vo, wo = oracle.sample_uniform(n, off, ntot, shift=2.0)
    np.testing.assert_allclose(d.get("v"), vo, rtol=1e-14, atol=1e-15)
    assert 0.0 <= d.get("v").min() and d.get("v").max() < 4.0
    vpm.initialize_(d, vpm.ShiftedNormalV((-5.0, 5.0), 2.0), offset=off, ntotal=ntot)
    xo, vo, wo = oracle.sample_maxwellian(n, off, ntot, xlo=-5.0, xhi=5.0, shift=2.0)
    np.testing.assert_allclose(d.get("x"), xo, rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(d.get("v"), vo, rtol=1e-13, atol=1e-14)
    assert abs(d.get("v").mean() - 2.0) < 0.1


def test_resample_spline_to_particles(vpm, oracle):
    """projection!(init::SplineDistribution, final::ParticleDistribution) (an empty TODO upstream,
    src/projections/distribution.jl:57-61): the device resampler against its CPU twin (which integrates by
    quadrature and inverts by bisection -- an independent route), slab-wise, with and without jitter, and the
    round trip particles -> spline -> particles -> spline at 2e6 particles."""
    nrm = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    vs = oracle.VSpace(-10.0, 10.0, 41, 4)
    src = vpm.ParticleDistribution(1, 1, 200000)
    vpm.initialize_(src, vpm.DoubleMaxwellian((-10.0, 10.0), 2.0))
    fs = vpm.projection(None, src, sd)
    c = fs.coefficients.copy()
    n, off, ntot = 30011, 7000, 50000
    for jitter in (False, True):
        d = vpm.ParticleDistribution(1, 1, n).set(np.arange(n, dtype=float), np.zeros(n), np.zeros(n))
        vpm.projection_(sd, d, offset=off, ntotal=ntot, jitter=jitter)
        vo, wo, mass = vs.resample(c, n, offset=off, Ntotal=ntot, jitter=jitter)
        x, v, w = d.get()
        np.testing.assert_array_equal(x, np.arange(n, dtype=float))          # positions untouched
        assert abs(d.resampled_mass - mass) < 1e-13 and np.allclose(w, wo, rtol=1e-13)
        assert nrm(v, vo) < 1e-11 and np.abs(v - vo).max() < 1e-8, (jitter, np.abs(v - vo).max())
    # round trip at size.  A projected (noisy) spline has small negative lobes in the tails, which the sampler
    # clips cell-wise: the re-projected coefficients agree up to that clipped part (independent of N) ...
    big = vpm.ParticleDistribution(1, 1, 2_000_000)
    vpm.projection_(sd, big)
    v = big.get("v")
    assert np.all(np.diff(v) >= 0.0) and abs(big.get("w").sum() - big.resampled_mass) < 1e-9
    sd2 = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    assert nrm(vpm.projection(None, big, sd2).coefficients, c) < 1e-3
    # ... and for a positive spline to the stratification error O(1/N): 2e-4 at 2e4 particles, 2e-6 at 2e6
    cpos = 0.1 * np.exp(-np.linspace(-10.0, 10.0, len(sd)) ** 2 / 8.0)
    vpm.projection_(sd, big, coefficients=cpos)
    assert nrm(vpm.projection(None, big, sd2).coefficients, cpos) < 1e-5
    with pytest.raises(vpm.VpmError):                                         # a spline with no positive mass
        sd3 = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
        sd3.mass_solve(-np.ones(len(sd3)))
        vpm.projection_(sd3, big)


def test_config1_vlasov_poisson_script(vpm, oracle):
    """scripts/vlasov_poisson.jl:6-30 with its shipped parameters, both field modes, against the oracle."""
    npart, nknot, order, tstep, tspan, domain = 10000, 16, 3, 0.1, (0.0, 20.0), (0.0, 1.0)
    dist = vpm.initialize_(vpm.ParticleDistribution(1, 1, npart), vpm.NormalDistribution(domain))
    x0, v0, w0 = dist.get()
    potential = vpm.Potential(vpm.PeriodicBasisBSplineKit(domain, order, nknot))
    model = vpm.VlasovPoisson(dist, potential)
    integrator = vpm.SplittingMethod(model, tspan, tstep)          # field="frozen": as shipped
    vpm.run_(integrator)
    xs = oracle.XSpace(domain[0], domain[1], order, nknot)
    xo, vo, _ = xs.strang_frozen(x0, v0, x0, w0, tstep, 200)
    x1, v1, _ = dist.get()
    nrm = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    # 200 steps: unwrapped positions reach |x| ~ 25 L, where one ulp of x is 6e-14 of a cell, and the per-kick
    # differences accumulate; the 1e-12 bar of the north star is for one step (tests/test_gpu_parity.py)
    assert nrm(x1, xo) < 1e-10 and nrm(v1, vo) < 1e-9
    # the physical (self-consistent) variant of the same configuration
    dist2 = vpm.ParticleDistribution(1, 1, npart).set(x0, v0, w0)
    integ2 = vpm.SplittingMethod(vpm.VlasovPoisson(dist2, potential), tspan, tstep, field="selfconsistent")
    vpm.run_(integ2, diag_mode=1)
    xo2, vo2, do2, _ = xs.strang_selfconsistent(x0, v0, w0, tstep, 200)
    x2, v2, _ = dist2.get()
    assert nrm(x2, xo2) < 1e-8 and nrm(v2, vo2) < 1e-7       # 200 steps of a chaotic N-body system
    np.testing.assert_allclose(integ2.diagnostics[:, 1], do2[:, 1], rtol=1e-8)
    np.testing.assert_allclose(integ2.diagnostics[:, 2], do2[:, 2], rtol=1e-6, atol=1e-9)
