"""Host-side logic of libvpm_b200 that needs no GPU: operator construction (Galerkin stencils, circulant
pseudo-inverse, clamped mass matrix and its banded Cholesky factor) against the oracle and scipy, and the
invariant-divisor index wrap used by the kernels."""
import ctypes as C

import numpy as np
import pytest
from scipy.interpolate import BSpline
from scipy.linalg import circulant


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import vpm_b200
    return vpm_b200._cabi.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("K,nh", [(2, 8), (3, 16), (4, 16), (5, 32), (6, 7), (4, 3)])
def test_periodic_galerkin_rows(lib, oracle, K, nh):
    lo, hi = -0.5, 3.1
    m, s, g = np.zeros(nh), np.zeros(nh), np.zeros(nh)
    assert lib.vpm_galerkin_periodic(lo, hi, K, nh, _p(m), _p(s), _p(g)) == 0
    M, S = oracle.XSpace(lo, hi, K, nh).matrices()        # Gauss-Legendre quadrature, independent method
    np.testing.assert_allclose(circulant(m), M, atol=1e-14)
    np.testing.assert_allclose(circulant(s), S, atol=1e-12)
    # pseudo-inverse: S G S = S, G 1 = 0 (zero-mean gauge), and it solves S phi = b for mean-free b
    G = circulant(g)
    np.testing.assert_allclose(S @ G @ S, S, atol=1e-11 * np.abs(S).max())
    assert np.abs(G.sum(axis=1)).max() < 1e-12 * np.abs(G).max()
    b = np.random.default_rng(0).standard_normal(nh)
    b -= b.mean()
    phi = G @ b
    np.testing.assert_allclose(phi, np.linalg.solve(S + 1.0 / nh, b), atol=1e-10 * np.abs(phi).max())


@pytest.mark.parametrize("K,nk,dirichlet", [(4, 41, 1), (4, 41, 0), (3, 10, 1), (5, 12, 1), (6, 9, 0), (2, 5, 1)])
def test_clamped_galerkin(lib, oracle, K, nk, dirichlet):
    lo, hi = -10.0, 10.0
    size = C.c_int()
    assert lib.vpm_galerkin_clamped(lo, hi, nk, K, dirichlet, C.byref(size), None, None) == 0
    nv = size.value
    assert nv == nk + K - 2 - 2 * dirichlet
    M, Lb = np.zeros((nv, nv)), np.zeros((nv, K))
    assert lib.vpm_galerkin_clamped(lo, hi, nk, K, dirichlet, C.byref(size), _p(M), _p(Lb)) == 0
    np.testing.assert_allclose(M, oracle.VSpace(lo, hi, nk, K, bool(dirichlet)).mass(), atol=1e-14)
    # scipy twin of the mass matrix
    br = np.linspace(lo, hi, nk)
    T = np.r_[[br[0]] * (K - 1), br, [br[-1]] * (K - 1)]
    xq, wq = np.polynomial.legendre.leggauss(K)
    Mref = np.zeros((nk + K - 2,) * 2)
    for a, b_ in zip(br[:-1], br[1:]):
        D = BSpline.design_matrix(0.5 * (a + b_) + 0.5 * (b_ - a) * xq, T, K - 1).toarray()
        Mref += D.T @ (D * (0.5 * (b_ - a) * wq)[:, None])
    if dirichlet:
        Mref = Mref[1:-1, 1:-1]
    np.testing.assert_allclose(M, Mref, atol=1e-14)
    # banded factor reproduces M = L L^T
    Lf = np.zeros((nv, nv))
    for i in range(nv):
        for k in range(K):
            if i - k >= 0:
                Lf[i, i - k] = Lb[i, k]
    np.testing.assert_allclose(Lf @ Lf.T, M, atol=1e-14)
    np.testing.assert_allclose(Lf, np.linalg.cholesky(M), atol=1e-13)


@pytest.mark.parametrize("d", [1, 2, 3, 7, 11, 16, 17, 41, 100, 127, 128, 1000, 4096, 65535, 65536, 1 << 20])
def test_index_wrap(lib, d):
    assert lib.vpm_selftest_wrap(d) == 0, lib.vpm_last_error()


def test_argument_errors_without_gpu(lib):
    assert lib.vpm_galerkin_periodic(0.0, 1.0, 9, 16, None, None, None) == -1
    assert b"bad arguments" in lib.vpm_last_error()
    assert lib.vpm_galerkin_periodic(1.0, 0.0, 4, 16, None, None, None) == -1
    assert lib.vpm_selftest_wrap(0) == -1


def test_plotting_forms_of_the_rhs_compose_the_operators(monkeypatch):
    """LB_rhs / CLB_rhs (the non-mutating "used for plotting" forms, lenard_bernstein.jl:37-44,
    lenard_bernstein_conservative.jl:53-64) and the *_GI_ argument order are pure compositions of the device
    operators; checked here with a host-side stand-in for the spline so that the composition itself is pinned
    without a GPU (the operators are pinned on the GPU in tests/test_gpu_parity.py)."""
    import vpm_b200 as vpm
    from vpm_b200 import api

    class FakeSplineDistribution:
        def evaluate(self, v, coefficients=None, derivative=False):
            v = np.asarray(v, dtype=float)
            g = np.exp(-v * v / 2) / np.sqrt(2 * np.pi)
            return -v * g if derivative else g

    sd = FakeSplineDistribution()
    fs = vpm.Spline(sd)
    v = np.linspace(-3, 3, 13)
    f, df = sd.evaluate(v), sd.evaluate(v, derivative=True)

    class M:
        pass
    model = M()
    model.ent = M()
    model.ent.dist = sd
    params = {"nu": 1.7, "idist": None, "fdist": sd, "model": model}
    np.testing.assert_allclose(vpm.LB_rhs(v, params, fs), -1.7 * (df + v * f), rtol=1e-15)   # = 0 for the unit Maxwellian
    assert np.abs(vpm.LB_rhs(v, params, fs)).max() < 1e-15
    m5 = np.array([f.sum(), (v * f).sum(), (v * v * f).sum(), df.sum(), (v * df).sum()])
    monkeypatch.setattr(api, "_moments", lambda dist, vp: tuple(m5))
    n, nu, ne, B1, B2 = m5[0], m5[1], m5[2], -m5[3], -m5[4]
    A1, A2 = (ne * B1 - nu * B2) / (n * ne - nu ** 2), -(nu * B1 - n * B2) / (n * ne - nu ** 2)
    got = vpm.CLB_rhs(v, params, fs)
    np.testing.assert_allclose(got, -1.7 * (df + (A1 + A2 * v) * f), rtol=1e-14, atol=1e-16)
    # the conservative form has no net momentum / energy input over the points the moments were taken on
    assert abs(got.sum()) < 1e-14 and abs((v * got).sum()) < 1e-14
    # GI argument order
    calls = []
    monkeypatch.setattr(api, "LB_rhs_", lambda vd, q, p, t=0.0: calls.append(("lb", t)) or vd)
    monkeypatch.setattr(api, "CLB_rhs_", lambda vd, q, p, t=0.0: calls.append(("clb", t)) or vd)
    out = np.zeros(3)
    assert vpm.LB_rhs_GI_(out, 0.25, np.ones(3), params) is out and vpm.CLB_rhs_GI_(out, 0.5, np.ones(3), params) is out
    assert calls == [("lb", 0.25), ("clb", 0.5)]


def test_ntime_is_a_ceiling_like_upstream():
    """GeometricEquations: ntime = Int(abs(div(tend - tbegin, tstep, RoundUp))) -- a tspan that is not a multiple of
    tstep gets the extra (partial-interval) step, and so does floating-point noise above an integer quotient"""
    from vpm_b200 import api
    assert api._ntime((0.0, 20.0), 0.1) == 200 and api._ntime((0.0, 500.0), 1e-2) == 50000 and api._ntime((0.0, 10.0), 0.1) == 100
    assert api._ntime((0.0, 1.05), 0.1) == 11                 # non-divisible: rounded UP (round() gave 10)
    assert api._ntime((0.0, 2.1), 0.3) == 8                   # 2.1 / 0.3 = 7.000000000000001: as upstream
    assert api._ntime((1.0, 0.0), 0.25) == 4
    for ns, dt in ((3, 0.1), (7, 0.013), (1, 0.29), (4, 0.05), (0, 0.1), (200, 0.1)):
        assert api._ntime(api.tspan_for(ns, dt), dt) == ns
