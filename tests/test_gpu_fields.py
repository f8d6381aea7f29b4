"""Host-array callback forms (vector fields, flows) and the legacy field functors, against the oracle.
Restates the reference's disabled test/electric_field_tests.jl:18-46 (PoissonField, ExternalField and
ScaledField(chi=1) produce identical fields for 100 bump-on-tail particles)."""
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


def test_vector_fields_and_flows(vpm, oracle):
    n, K, nh, L = 5000, 3, 16, 1.0
    rng = np.random.default_rng(4)
    x0, v0, w = rng.uniform(0, L, n), rng.standard_normal(n), np.full(n, 1.0 / n)
    dist = vpm.ParticleDistribution(1, 1, n).set(x0, v0, w)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), K, nh))
    model = vpm.VlasovPoisson(dist, pot)
    params = {"phi": pot, "model": model}
    xs = oracle.XSpace(0.0, L, K, nh)
    phi = xs.poisson_solve(xs.deposit(x0, w))          # field of model.distribution (frozen)
    z = np.vstack([x0 + 0.3 * v0, v0 * 0.5])           # integrator state differs from the distribution
    zdot = np.zeros_like(z)
    vpm.lorentz_force_(zdot, 0.0, z, params)
    np.testing.assert_array_equal(zdot[0], z[1])
    assert nrm(zdot[1], -xs.eval(phi, z[0], 1)) < 1e-11
    vpm.v_acceleration_(zdot, 0.0, z, params)
    assert np.all(zdot[0] == 0) and nrm(zdot[1], -xs.eval(phi, z[0], 1)) < 1e-11
    vpm.v_advection_(zdot, 0.0, z, params)
    np.testing.assert_array_equal(zdot[0], z[1])
    zn = np.zeros_like(z)
    vpm.s_advection_host_(zn, 0.25, z, 0.1, params)
    assert nrm(zn[0], oracle.push_drift(z[0], z[1], 0.15)) < 1e-15
    np.testing.assert_array_equal(zn[1], z[1])
    vpm.s_acceleration_host_(zn, 0.25, z, 0.1, params)
    np.testing.assert_array_equal(zn[0], z[0])
    assert nrm(zn[1], xs.push_kick(phi, z[0], z[1], 0.15)) < 1e-12


def test_legacy_field_functors(vpm, oracle):
    n, K, nh = 100, 4, 16
    bot = vpm.BumpOnTail()
    x, v, w = oracle.sample_bump_on_tail(n)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, bot.L), K, nh))
    xs = oracle.XSpace(0.0, bot.L, K, nh)
    f1 = vpm.PoissonField(pot)
    e1 = f1(np.zeros(n), x, w, 0.0)
    phi = xs.poisson_solve(xs.deposit(x, w))
    assert nrm(e1, -xs.eval(phi, x, 1)) < 1e-11
    assert abs(f1.energy() - xs.field_energy(phi)) < 1e-11 * xs.field_energy(phi)
    # ExternalField with the same coefficients in column ts reproduces the Poisson field exactly
    nt = 4
    coeffs = np.zeros((nh, nt + 1))
    coeffs[:, 2] = f1.coefficients()
    f2 = vpm.ExternalField(pot, coeffs, 0.1)
    e2 = f2(np.zeros(n), x, w, 0.2)
    np.testing.assert_array_equal(e2, e1)
    f3 = vpm.ScaledPoissonField(pot, 1.0)
    e3 = f3(np.zeros(n), x, w, 0.0)
    np.testing.assert_array_equal(e3, e1)
    f4 = vpm.ScaledPoissonField(pot, 2.0)
    e4 = f4(np.zeros(n), x, w, 0.0)
    assert nrm(e4, e1 / 4) < 1e-15 and abs(f4.energy() - f1.energy() / 4) < 1e-15


def test_plain_c_driver_runs():
    """The ABI is usable from plain C without Python or torch (examples/vp_bump_on_tail.c)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "vp_bump_on_tail")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "vp_bump_on_tail.c"),
                           "-o", exe, "-L" + os.path.join(root, "vlasovparticlemethods.jl_b200", "lib"), "-lvpm_b200",
                           "-Wl,-rpath," + os.path.join(root, "vlasovparticlemethods.jl_b200", "lib"), "-lm"])
    out = subprocess.run([exe, "2000000", "100"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    last = out.stdout.strip().splitlines()[-1]
    assert "particle-steps/s" in last
    drift = float(last.split("energy drift")[1])
    assert drift < 5e-3
    # the same driver with trajectory output: run!(method, h5file) from plain C
    import tempfile
    import h5mini
    with tempfile.TemporaryDirectory() as tmp:
        h5 = os.path.join(tmp, "bot.h5")
        out = subprocess.run([exe, "200000", "100", h5, "40"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        assert "4 frames" in out.stdout.strip().splitlines()[-1]
        f = h5mini.File(h5)
        assert f.datasets["z"].shape == (4, 200000, 2)
        t = f.read("t")
        assert abs(t[-1] - 10.0) < 1e-12 and abs(t[1] - 4.0) < 1e-12
