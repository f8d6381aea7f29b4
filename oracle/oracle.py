"""ctypes binding of the CPU oracle (oracle/vpm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED (see the header of vpm_oracle.c): pinned by analytic known-answer tests
and a scipy twin, not by reference golden vectors (the reference has none).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvpm_oracle.so")
_lib = None

_D = C.POINTER(C.c_double)


def build(force=False):
    """Compile oracle/libvpm_oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "vpm_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libvpm_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _dp(a):
    return a.ctypes.data_as(_D) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        i64, i32, f64, u64, u32 = C.c_int64, C.c_int, C.c_double, C.c_uint64, C.c_uint32
        sig = {
            "vpo_set_threads": (None, [i32]),
            "vpo_max_threads": (i32, []),
            "vpo_xspace_create": (vp, [f64, f64, i32, i32]),
            "vpo_xspace_destroy": (None, [vp]),
            "vpo_xspace_matrices": (None, [vp, _D, _D]),
            "vpo_xbasis": (i32, [vp, f64, _D]),
            "vpo_deposit_x": (None, [vp, i64, _D, _D, _D]),
            "vpo_poisson_solve": (None, [vp, _D, _D]),
            "vpo_mass_solve_x": (None, [vp, _D, _D]),
            "vpo_xeval": (f64, [vp, _D, f64, i32]),
            "vpo_xeval_many": (None, [vp, _D, i64, _D, i32, _D]),
            "vpo_field_energy": (f64, [vp, _D]),
            "vpo_push_drift": (None, [i64, _D, _D, f64]),
            "vpo_push_kick": (None, [vp, _D, i64, _D, _D, f64, f64]),
            "vpo_vp_strang_selfconsistent": (None, [vp, i64, _D, _D, _D, f64, f64, i32, _D, _D]),
            "vpo_vp_strang_frozen": (None, [vp, i64, _D, _D, i64, _D, _D, f64, i32, _D]),
            "vpo_vspace_create": (vp, [f64, f64, i32, i32, i32]),
            "vpo_vspace_destroy": (None, [vp]),
            "vpo_vspace_size": (i32, [vp]),
            "vpo_vspace_mass": (None, [vp, _D]),
            "vpo_vbasis": (i32, [vp, f64, _D, i32]),
            "vpo_deposit_v": (None, [vp, i64, _D, _D, _D]),
            "vpo_mass_solve_v": (None, [vp, _D, _D]),
            "vpo_project_v": (None, [vp, i64, _D, _D, _D]),
            "vpo_veval": (f64, [vp, _D, f64, i32]),
            "vpo_veval_many": (None, [vp, _D, i64, _D, i32, _D]),
            "vpo_moments": (None, [vp, _D, i64, _D, _D]),
            "vpo_clb_coefficients": (None, [_D, _D]),
            "vpo_lb_rhs": (None, [vp, i64, _D, _D, f64, i32, _D, _D, _D]),
            "vpo_lb_rk438": (None, [vp, i64, _D, _D, f64, f64, i32, i32, _D]),
            "vpo_entropy_v": (f64, [vp, _D, i64, _D, _D, f64, C.POINTER(C.c_double)]),
            "vpo_lb_rk438_entropy": (None, [vp, i64, _D, _D, f64, f64, i32, i32, f64, _D, _D, _D]),
            "vpo_uniform": (f64, [u64, u64, u32]),
            "vpo_norminv": (f64, [f64]),
            "vpo_sample_bump_on_tail": (None, [i64, i64, i64, u64, f64, f64, f64, f64, f64, _D, _D, _D]),
            "vpo_sample_normal": (f64, [i64, i64, i64, u64, f64, f64, f64, _D, _D, _D]),
            "vpo_sample_maxwellian": (None, [i64, i64, i64, u64, f64, f64, f64, i32, f64, _D, _D, _D]),
            "vpo_resample_v": (f64, [vp, _D, i64, i64, i64, u64, i32, _D, _D]),
            "vpo_sample_uniform": (None, [i64, i64, i64, u64, f64, f64, f64, f64, f64, f64, _D, _D, _D]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def set_threads(n):
    lib().vpo_set_threads(int(n))


def max_threads():
    return int(lib().vpo_max_threads())


class XSpace:
    """Periodic uniform B-spline space (PeriodicBasisBSplineKit(domain, order, n_basis))."""

    def __init__(self, lo, hi, order, nh):
        self.lo, self.hi, self.K, self.nh = float(lo), float(hi), int(order), int(nh)
        self.h = (self.hi - self.lo) / self.nh
        self._h = lib().vpo_xspace_create(self.lo, self.hi, self.K, self.nh)
        if not self._h:
            raise ValueError("bad x-space parameters")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vpo_xspace_destroy(self._h)
            self._h = None

    def matrices(self):
        M = np.zeros((self.nh, self.nh))
        S = np.zeros((self.nh, self.nh))
        lib().vpo_xspace_matrices(self._h, _dp(M), _dp(S))
        return M, S

    def basis(self, x):
        b = np.zeros(self.K)
        c = lib().vpo_xbasis(self._h, float(x), _dp(b))
        return c, b

    def deposit(self, x, w):
        x, w = _f64(x), _f64(w)
        rhs = np.zeros(self.nh)
        lib().vpo_deposit_x(self._h, x.size, _dp(x), _dp(w), _dp(rhs))
        return rhs

    def poisson_solve(self, rhs):
        rhs = _f64(rhs)
        phi = np.zeros(self.nh)
        lib().vpo_poisson_solve(self._h, _dp(rhs), _dp(phi))
        return phi

    def mass_solve(self, rhs):
        rhs = _f64(rhs)
        out = np.zeros(self.nh)
        lib().vpo_mass_solve_x(self._h, _dp(rhs), _dp(out))
        return out

    def eval(self, coef, x, deriv=0):
        coef, x = _f64(coef), _f64(np.atleast_1d(x))
        out = np.zeros(x.size)
        lib().vpo_xeval_many(self._h, _dp(coef), x.size, _dp(x), int(deriv), _dp(out))
        return out

    def field_energy(self, phi):
        phi = _f64(phi)
        return float(lib().vpo_field_energy(self._h, _dp(phi)))

    def push_kick(self, phi, x, v, tau, scale=1.0):
        phi, x = _f64(phi), _f64(x)
        v = _f64(v).copy()
        lib().vpo_push_kick(self._h, _dp(phi), x.size, _dp(x), _dp(v), float(tau), float(scale))
        return v

    def strang_selfconsistent(self, x, v, w, dt, nsteps, chi=1.0, diag=True):
        x, v, w = _f64(x).copy(), _f64(v).copy(), _f64(w)
        d = np.zeros((nsteps + 1, 3)) if diag else None
        phi = np.zeros(self.nh)
        lib().vpo_vp_strang_selfconsistent(self._h, x.size, _dp(x), _dp(v), _dp(w), float(dt), float(chi),
                                           int(nsteps), _dp(d), _dp(phi))
        return x, v, d, phi

    def strang_frozen(self, x, v, xdep, wdep, dt, nsteps):
        x, v = _f64(x).copy(), _f64(v).copy()
        xdep, wdep = _f64(xdep), _f64(wdep)
        phi = np.zeros(self.nh)
        lib().vpo_vp_strang_frozen(self._h, x.size, _dp(x), _dp(v), xdep.size, _dp(xdep), _dp(wdep),
                                   float(dt), int(nsteps), _dp(phi))
        return x, v, phi


def push_drift(x, v, tau):
    x = _f64(x).copy()
    v = _f64(v)
    lib().vpo_push_drift(x.size, _dp(x), _dp(v), float(tau))
    return x


class VSpace:
    """Clamped (optionally Dirichlet-recombined) spline space of SplineDistribution."""

    def __init__(self, lo, hi, nknots, order, dirichlet=True):
        self.lo, self.hi, self.nknots, self.K = float(lo), float(hi), int(nknots), int(order)
        self.dirichlet = bool(dirichlet)
        self._h = lib().vpo_vspace_create(self.lo, self.hi, self.nknots, self.K, int(self.dirichlet))
        if not self._h:
            raise ValueError("bad v-space parameters")
        self.nv = int(lib().vpo_vspace_size(self._h))
        self.h = (self.hi - self.lo) / (self.nknots - 1)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vpo_vspace_destroy(self._h)
            self._h = None

    def mass(self):
        M = np.zeros((self.nv, self.nv))
        lib().vpo_vspace_mass(self._h, _dp(M))
        return M

    def basis(self, v, deriv=0):
        b = np.zeros(self.K)
        c = lib().vpo_vbasis(self._h, float(v), _dp(b), int(deriv))
        return c, b

    def deposit(self, v, w):
        v, w = _f64(v), _f64(w)
        rhs = np.zeros(self.nv)
        lib().vpo_deposit_v(self._h, v.size, _dp(v), _dp(w), _dp(rhs))
        return rhs

    def mass_solve(self, rhs):
        rhs = _f64(rhs)
        c = np.zeros(self.nv)
        lib().vpo_mass_solve_v(self._h, _dp(rhs), _dp(c))
        return c

    def project(self, v, w):
        v, w = _f64(v), _f64(w)
        c = np.zeros(self.nv)
        lib().vpo_project_v(self._h, v.size, _dp(v), _dp(w), _dp(c))
        return c

    def eval(self, coef, v, deriv=0):
        coef, v = _f64(coef), _f64(np.atleast_1d(v))
        out = np.zeros(v.size)
        lib().vpo_veval_many(self._h, _dp(coef), v.size, _dp(v), int(deriv), _dp(out))
        return out

    def moments(self, coef, v):
        coef, v = _f64(coef), _f64(v)
        out = np.zeros(5)
        lib().vpo_moments(self._h, _dp(coef), v.size, _dp(v), _dp(out))
        return out

    def lb_rhs(self, v, w, nu=1.0, conservative=False):
        v, w = _f64(v), _f64(w)
        vdot = np.zeros(v.size)
        coef = np.zeros(self.nv)
        A = np.zeros(2)
        lib().vpo_lb_rhs(self._h, v.size, _dp(v), _dp(w), float(nu), int(conservative), _dp(vdot), _dp(coef), _dp(A))
        return vdot, coef, A

    def resample(self, coef, N, offset=0, Ntotal=None, seed=0x5EED0001, jitter=False):
        """spline -> particles by stratified inverse-CDF sampling (quadrature + bisection); returns v, w, mass"""
        coef = _f64(coef)
        Ntotal = N if Ntotal is None else Ntotal
        v, w = np.empty(N), np.empty(N)
        mass = lib().vpo_resample_v(self._h, _dp(coef), N, offset, Ntotal, seed, int(jitter), _dp(v), _dp(w))
        return v, w, mass

    def entropy(self, coef, v, w, f_floor=1e-14):
        """S = -sum w ln max(f_s(v), f_floor) (non-reference diagnostic, see vpo_entropy_v); returns (S, n_floored)"""
        coef, v, w = _f64(coef), _f64(v), _f64(w)
        nf = C.c_double()
        S = lib().vpo_entropy_v(self._h, _dp(coef), v.size, _dp(v), _dp(w), float(f_floor), C.byref(nf))
        return float(S), int(nf.value)

    def rk438_entropy(self, v, w, nu, dt, nsteps, conservative=False, f_floor=1e-14):
        v, w = _f64(v).copy(), _f64(w)
        d, e, nf = np.zeros((nsteps + 1, 2)), np.zeros(nsteps + 1), np.zeros(nsteps + 1)
        lib().vpo_lb_rk438_entropy(self._h, v.size, _dp(v), _dp(w), float(nu), float(dt), int(conservative), int(nsteps),
                                   float(f_floor), _dp(d), _dp(e), _dp(nf))
        return v, d, e, nf

    def rk438(self, v, w, nu, dt, nsteps, conservative=False, diag=True):
        v, w = _f64(v).copy(), _f64(w)
        d = np.zeros((nsteps + 1, 2)) if diag else None
        lib().vpo_lb_rk438(self._h, v.size, _dp(v), _dp(w), float(nu), float(dt), int(conservative), int(nsteps), _dp(d))
        return v, d


def clb_coefficients(m5):
    m5 = _f64(m5)
    A = np.zeros(2)
    lib().vpo_clb_coefficients(_dp(m5), _dp(A))
    return A


def uniform(seed, idx, stream):
    return float(lib().vpo_uniform(int(seed), int(idx), int(stream)))


def norminv(p):
    return float(lib().vpo_norminv(float(p)))


def sample_bump_on_tail(N, offset=0, Ntotal=None, seed=0x5EED0001, eps=0.03, kappa=0.3, alpha=0.1, sigma=0.5, v0=4.5):
    Ntotal = N if Ntotal is None else Ntotal
    x, v, w = np.zeros(N), np.zeros(N), np.zeros(N)
    lib().vpo_sample_bump_on_tail(N, offset, Ntotal, seed, eps, kappa, alpha, sigma, v0, _dp(x), _dp(v), _dp(w))
    return x, v, w


def sample_normal(N, offset=0, Ntotal=None, seed=0x5EED0001, xlo=0.0, xhi=1.0, xmax=0.0):
    Ntotal = N if Ntotal is None else Ntotal
    x, v, w = np.zeros(N), np.zeros(N), np.zeros(N)
    used = lib().vpo_sample_normal(N, offset, Ntotal, seed, xlo, xhi, xmax, _dp(x), _dp(v), _dp(w))
    return x, v, w, used


def sample_uniform(N, offset=0, Ntotal=None, seed=0x5EED0001, xlo=0.0, xhi=1.0, vlo=-2.0, vhi=2.0, shift=0.0, wnum=1.0):
    Ntotal = N if Ntotal is None else Ntotal
    x, v, w = np.empty(N), np.empty(N), np.empty(N)
    lib().vpo_sample_uniform(N, offset, Ntotal, seed, xlo, xhi, vlo, vhi, shift, wnum, _dp(x), _dp(v), _dp(w))
    return x, v, w


def sample_maxwellian(N, offset=0, Ntotal=None, seed=0x5EED0001, xlo=0.0, xhi=1.0, shift=0.0, doubled=False, wnum=1.0):
    Ntotal = N if Ntotal is None else Ntotal
    x, v, w = np.zeros(N), np.zeros(N), np.zeros(N)
    lib().vpo_sample_maxwellian(N, offset, Ntotal, seed, xlo, xhi, shift, int(doubled), wnum, _dp(x), _dp(v), _dp(w))
    return x, v, w
