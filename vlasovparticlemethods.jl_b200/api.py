"""Host-side mirror of the VlasovMethods.jl interface for the particle hot path, over the C ABI.

Julia is absent from this image, so the host layer the north star asks for in Julia is
written in Python with the reference's names, argument meaning and error behaviour; the same calls
as `ccall`s are in julia/VPMB200.jl.  Julia's `f!` becomes `f_` here.  Nothing in this module
computes: every operation is one call into libvpm_b200.so (no CPU fallback).

Reference files mirrored (paths relative to the reference checkout):
  ParticleDistribution      src/distributions/particle_distribution.jl:2-24
  SplineDistribution        src/distributions/spline_distribution.jl:1-36
  Potential / projection!   scripts/vlasov_poisson.jl:21, src/projections/potential.jl:2-22
  projection, moments       src/projections/distribution.jl:35-55, src/projections/density.jl:6-52
  VlasovPoisson, flows      src/models/vlasov_poisson.jl:2-89
  LenardBernstein (+cons.)  src/models/lenard_bernstein.jl, src/models/lenard_bernstein_conservative.jl
  SplittingMethod, run!     src/methods/splitting.jl:2-52 ; GeometricIntegrator: src/methods/geometric_integrator.jl
  examples / initialize!    src/examples/{normal,bumpontail,doublemaxwellian,uniform}.jl
"""
import ctypes as C
import math
import os

import numpy as np

from . import _cabi
from ._cabi import VpmError, check

_vp = C.c_void_p
DEFAULT_SEED = 0x5EED0001   # seed of the counter-based device samplers


def _lib():
    return _cabi.lib()


def _hp(a):
    """host pointer of a contiguous float64 array (or None)"""
    return None if a is None else a.ctypes.data_as(_vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ------------------------------------------------------------------------------------------------
# context
# ------------------------------------------------------------------------------------------------
class Context:
    """One per GPU (vpm_ctx).  `stream`: raw cudaStream_t (int) to enqueue on, or None for a private one."""

    def __init__(self, device=0, stream=None):
        h = _vp()
        check(_lib().vpm_ctx_create(int(device), _vp(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            _lib().vpm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(_lib().vpm_sync(self._h))

    @property
    def launches(self):
        return int(_lib().vpm_launch_count(self._h))

    def device_info(self):
        sm, sh, tot = C.c_int(), C.c_int64(), C.c_int64()
        check(_lib().vpm_device_info(self._h, C.byref(sm), C.byref(sh), C.byref(tot)))
        return {"sm_count": sm.value, "smem_optin": sh.value, "total_mem": tot.value}

    # multi-GPU: particle slabs + all-reduce of the coefficient vectors
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        check(_lib().vpm_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(_lib().vpm_comm_init(self._h, int(nranks), int(rank), buf))

    def comm_destroy(self):
        check(_lib().vpm_comm_destroy(self._h))

    # fused peer-memory all-reduce over NVLink (preferred): handles are exchanged by the host
    def p2p_prepare(self):
        buf = (C.c_char * 64)()
        check(_lib().vpm_p2p_prepare(self._h, buf))
        return bytes(buf)

    def p2p_attach(self, nranks, rank, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * nranks
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        check(_lib().vpm_p2p_attach(self._h, int(nranks), int(rank), buf))

    def p2p_detach(self):
        check(_lib().vpm_p2p_detach(self._h))

    def p2p_check(self):
        seq = C.c_uint64()
        check(_lib().vpm_p2p_error(self._h, C.byref(seq)))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


class DeviceVector:
    """Plain device buffer of doubles (vpm_dev_alloc) for integrator states handed to the operators."""

    def __init__(self, ctx, n, data=None):
        self.ctx, self.n = ctx, int(n)
        p = _vp()
        check(_lib().vpm_dev_alloc(ctx._h, self.n, C.byref(p)))
        self.ptr = p
        if data is not None:
            self.upload(data)

    def upload(self, a):
        a = _f64(a).ravel()
        assert a.size == self.n
        check(_lib().vpm_memcpy_h2d(self.ctx._h, self.ptr, _hp(a), self.n))

    def download(self):
        out = np.empty(self.n)
        check(_lib().vpm_memcpy_d2h(self.ctx._h, _hp(out), self.ptr, self.n))
        return out

    def free(self):
        if getattr(self, "ptr", None) and self.ctx._h:
            _lib().vpm_dev_free(self.ctx._h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# distributions
# ------------------------------------------------------------------------------------------------
class DistributionFunction:
    """abstract type DistributionFunction{XD,VD} (src/distributions/distribution.jl:2)"""


class _ParticleViews:
    """`.particles` of the reference ParticleDistribution: named row views x, v, w, z of the
    (xdim+vdim+1) x N matrix (particle_distribution.jl:11-17).  Reads download, writes upload."""

    def __init__(self, dist):
        self._d = dist

    def __len__(self):
        return self._d.npart

    @property
    def x(self):
        return self._d.get("x")[None, :]

    @x.setter
    def x(self, val):
        self._d.set(x=np.asarray(val).reshape(-1))

    @property
    def v(self):
        return self._d.get("v")[None, :]

    @v.setter
    def v(self, val):
        self._d.set(v=np.asarray(val).reshape(-1))

    @property
    def w(self):
        return self._d.get("w")[None, :]

    @w.setter
    def w(self, val):
        self._d.set(w=np.asarray(val).reshape(-1))

    @property
    def z(self):
        return self._d.download_aos(2)

    @z.setter
    def z(self, val):
        self._d.upload_aos(np.asarray(val))


class ParticleDistribution(DistributionFunction):
    """ParticleDistribution(xdim, vdim, npart): device-resident SoA particle store."""

    def __init__(self, xdim, vdim, npart, ctx=None):
        if xdim != 1 or vdim != 1:
            raise ValueError("the B200 hot path is 1D1V (VlasovPoisson{1,1}, LenardBernstein{1,1})")
        self.xdim, self.vdim, self.npart = 1, 1, int(npart)
        self.ctx = ctx or default_context()
        h = _vp()
        check(_lib().vpm_particles_create(self.ctx._h, self.npart, C.byref(h)))
        self._h = h
        self.particles = _ParticleViews(self)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.ctx._h:
                _lib().vpm_particles_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def size(self):
        return self.npart

    def __len__(self):
        return self.npart

    def ptrs(self, writable=False):
        """raw device pointers (x, v, w).  writable=True hands out the mutable pointers and thereby ends a uniform-weight
        declaration (the caller may rewrite w); the read-only form, used by every operator of this module, does not."""
        x, v, w = _vp(), _vp(), _vp()
        fn = _lib().vpm_particles_ptrs if writable else _lib().vpm_particles_ptrs_const
        check(fn(self._h, C.byref(x), C.byref(v), C.byref(w)))
        return x, v, w

    def set(self, x=None, v=None, w=None):
        x = None if x is None else _f64(x)
        v = None if v is None else _f64(v)
        w = None if w is None else _f64(w)
        for a in (x, v, w):
            if a is not None and a.size != self.npart:
                raise ValueError("array length does not match the number of particles")
        check(_lib().vpm_particles_upload_soa(self._h, _hp(x), _hp(v), _hp(w)))
        return self

    def get(self, which=None):
        if which is not None:
            out = np.empty(self.npart)
            args = {"x": (_hp(out), None, None), "v": (None, _hp(out), None), "w": (None, None, _hp(out))}[which]
            check(_lib().vpm_particles_download_soa(self._h, *args))
            return out
        x, v, w = np.empty(self.npart), np.empty(self.npart), np.empty(self.npart)
        check(_lib().vpm_particles_download_soa(self._h, _hp(x), _hp(v), _hp(w)))
        return x, v, w

    def set_uniform_weight(self, w):
        """all particles carry weight w (as every reference sampler produces): the steppers skip the w stream"""
        check(_lib().vpm_particles_set_uniform_weight(self._h, float(w)))
        return self

    def upload_aos(self, z):
        """z: (ld, N) Julia column-major matrix == C-order (N, ld) array; rows x, v[, w]."""
        z = np.asarray(z, dtype=np.float64)
        ld = z.shape[0]
        if z.ndim != 2 or ld not in (2, 3) or z.shape[1] != self.npart:
            raise ValueError("expected a (2|3) x N matrix")
        zz = np.ascontiguousarray(z.T)  # memory order of Julia's column-major ld x N
        check(_lib().vpm_particles_upload_aos(self._h, _hp(zz), ld))
        return self

    def download_aos(self, ld=3):
        zz = np.empty((self.npart, ld))
        check(_lib().vpm_particles_download_aos(self._h, _hp(zz), ld))
        return zz.T


class SplineDistribution(DistributionFunction):
    """SplineDistribution(xdim, vdim, nknots, order, domain, bc=:Dirichlet)"""

    def __init__(self, xdim, vdim, nknots, order, domain, bc="Dirichlet", ctx=None):
        self.ctx = ctx or default_context()
        self.nknots, self.order, self.domain = int(nknots), int(order), (float(domain[0]), float(domain[1]))
        self.bc = str(bc).lstrip(":")
        h = _vp()
        check(_lib().vpm_vspace_create(self.ctx._h, self.domain[0], self.domain[1], self.nknots, self.order,
                                       1 if self.bc == "Dirichlet" else 0, C.byref(h)))
        self._h = h
        self.nbasis = int(_lib().vpm_vspace_size(h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.ctx._h:
                _lib().vpm_vspace_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def size(self):
        return (self.nbasis,)

    def __len__(self):
        return self.nbasis

    @property
    def coefficients(self):
        c = np.empty(self.nbasis)
        check(_lib().vpm_vspace_get(self._h, None, _hp(c)))
        return c

    @coefficients.setter
    def coefficients(self, c):
        self._set_coefficients(_f64(c))

    def _set_coefficients(self, c):
        # uploading coefficients == evaluating with an explicit coefficient vector once (rebuilds the tables)
        dummy = DeviceVector(self.ctx, 2, np.zeros(2))
        check(_lib().vpm_gather_v(self._h, _hp(c), dummy.ptr, 0, None, None))
        dummy.free()

    @property
    def rhs(self):
        r = np.empty(self.nbasis)
        check(_lib().vpm_vspace_get(self._h, _hp(r), None))
        return r

    @property
    def mass_matrix(self):
        M = np.empty((self.nbasis, self.nbasis))
        check(_lib().vpm_vspace_mass(self._h, _hp(M)))
        return M

    def mass_solve(self, rhs):
        """mass_fact \\ rhs (ldiv!, src/projections/distribution.jl:52)"""
        rhs = _f64(rhs)
        c = np.empty(self.nbasis)
        check(_lib().vpm_mass_solve_v(self._h, _hp(rhs), _hp(c)))
        return c

    def evaluate(self, v, coefficients=None, derivative=False):
        """spline.(v) / (Derivative(1)*spline).(v) for a host array v"""
        v = _f64(np.atleast_1d(v)).ravel()
        dv = DeviceVector(self.ctx, v.size, v)
        out = DeviceVector(self.ctx, v.size)
        c = None if coefficients is None else _f64(coefficients)
        if derivative:
            check(_lib().vpm_gather_v(self._h, _hp(c), dv.ptr, v.size, None, out.ptr))
        else:
            check(_lib().vpm_gather_v(self._h, _hp(c), dv.ptr, v.size, out.ptr, None))
        r = out.download()
        dv.free(); out.free()
        return r

    def __call__(self, v):
        return self.evaluate(v)


class Spline:
    """Result of `projection`: aliases the SplineDistribution's coefficients (spline_distribution.jl:9)."""

    def __init__(self, sdist, derivative=False):
        self.sdist, self.derivative = sdist, derivative

    @property
    def coefficients(self):
        return self.sdist.coefficients

    def __call__(self, v):
        return self.sdist.evaluate(v, derivative=self.derivative)


def Derivative(n=1):
    class _D:
        order = n

        def __mul__(self, spline):
            if n != 1 or not isinstance(spline, Spline) or spline.derivative:
                raise VpmError(-5, "only Derivative(1) * Spline is on the hot path")
            return Spline(spline.sdist, derivative=True)
    return _D()


# ------------------------------------------------------------------------------------------------
# potential (PoissonSolvers.Potential over a periodic B-spline basis)
# ------------------------------------------------------------------------------------------------
class PeriodicBasisBSplineKit:
    """PeriodicBasisBSplineKit(domain, order, n): n = number of periodic basis functions
    (whether PoissonSolvers' `nknot` means n or n+1 is unpinned; here it is n)."""

    def __init__(self, domain, order, n):
        self.domain, self.order, self.n = (float(domain[0]), float(domain[1])), int(order), int(n)


class Potential:
    def __init__(self, basis, ctx=None):
        self.ctx = ctx or default_context()
        self.basis = basis
        h = _vp()
        check(_lib().vpm_xspace_create(self.ctx._h, basis.domain[0], basis.domain[1], basis.order, basis.n, C.byref(h)))
        self._h = h
        self.n = basis.n

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.ctx._h:
                _lib().vpm_xspace_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def rhs(self):
        r = np.empty(self.n)
        check(_lib().vpm_xspace_get(self._h, _hp(r), None))
        return r

    @property
    def coefficients(self):
        c = np.empty(self.n)
        check(_lib().vpm_xspace_get(self._h, None, _hp(c)))
        return c

    def stencils(self):
        K = self.basis.order
        m, s = np.empty(2 * K - 1), np.empty(2 * K - 1)
        check(_lib().vpm_xspace_stencils(self._h, _hp(m), _hp(s)))
        return m, s

    def mass_solve(self, rhs):
        """potential.solver.Mfac \\ rhs (test/projections_tests.jl:27)"""
        rhs = _f64(rhs)
        out = np.empty(self.n)
        check(_lib().vpm_mass_solve_x(self._h, _hp(rhs), _hp(out)))
        return out

    def solve(self, rhs):
        rhs = _f64(rhs)
        phi = np.empty(self.n)
        check(_lib().vpm_poisson_solve(self._h, _hp(rhs), _hp(phi)))
        return phi

    def energy(self, phi=None):
        phi = self.coefficients if phi is None else _f64(phi)
        e = C.c_double()
        check(_lib().vpm_field_energy(self._h, _hp(phi), C.byref(e)))
        return e.value

    def evaluate(self, x, derivative=0, coefficients=None):
        """phi(x) / phi(x, Derivative(1)) for host positions x"""
        x = _f64(np.atleast_1d(x)).ravel()
        dx = DeviceVector(self.ctx, x.size, x)
        out = DeviceVector(self.ctx, x.size)
        c = None if coefficients is None else _f64(coefficients)
        check(_lib().vpm_gather_x(self._h, _hp(c), dx.ptr, x.size, int(derivative), out.ptr))
        r = out.download()
        dx.free(); out.free()
        return r

    def __call__(self, x, derivative=None):
        d = 0 if derivative is None else getattr(derivative, "order", int(derivative))
        r = self.evaluate(x, d)
        return r if np.ndim(x) else float(r[0])


def update_(potential):
    """PoissonSolvers.update!(potential): rhs -> coefficients (call site vlasov_poisson.jl:14)"""
    check(_lib().vpm_poisson_solve(potential._h, None, None))
    return potential


def projection_(init, final, seed=DEFAULT_SEED, offset=0, ntotal=None, jitter=False, coefficients=None):
    """projection!(a, b), dispatched on the argument types like the reference's methods:

    projection!(potential::Potential, distribution::ParticleDistribution)     src/projections/potential.jl:2-22
        charge deposit into potential.rhs;
    projection!(init::ParticleDistribution, final::SplineDistribution)        src/projections/distribution.jl:7-32
        (commented out upstream but called by update_entropy!, lenard_bernstein.jl:15-17): deposit + mass solve;
    projection!(init::SplineDistribution, final::ParticleDistribution)        src/projections/distribution.jl:57-61
        (an empty TODO upstream): resample the particle velocities from the spline by stratified inverse-CDF
        sampling, equal weights summing to the integral of the spline; x is left untouched (`coefficients`:
        spline coefficients to sample from instead of those of the last projection);
    same-type pairs                                                            src/projections/distribution.jl:64-71
        do nothing, as upstream.
    """
    if isinstance(init, Potential):
        x, v, w = final.ptrs()
        check(_lib().vpm_deposit_x(init._h, x, w, final.npart, None))
        return init
    if isinstance(init, ParticleDistribution) and isinstance(final, SplineDistribution):
        projection(None, init, final)
        return final
    if isinstance(init, SplineDistribution) and isinstance(final, ParticleDistribution):
        mass = C.c_double()
        coef = None if coefficients is None else _f64(coefficients)
        if coef is not None and coef.size != len(init):
            raise ValueError("coefficients must have one entry per basis function")
        check(_lib().vpm_resample_v(init._h, _hp(coef), final._h, int(offset), int(final.npart if ntotal is None else ntotal),
                                    int(seed), int(bool(jitter)), C.byref(mass)))
        final.resampled_mass = mass.value
        return final
    if type(init) is type(final):
        return final
    raise TypeError(f"no projection_ method for ({type(init).__name__}, {type(final).__name__})")


def projection(velocities, dist, final_dist):
    """projection(velocities, dist, final_dist) -> Spline: src/projections/distribution.jl:35-55"""
    x, v, w = dist.ptrs()
    if velocities is None:
        check(_lib().vpm_project_v(final_dist._h, v, w, dist.npart, None))
    else:
        vel = _f64(velocities).ravel()
        if vel.size != dist.npart:
            raise ValueError("velocities must have one entry per particle")
        dv = DeviceVector(dist.ctx, vel.size, vel)
        check(_lib().vpm_project_v(final_dist._h, dv.ptr, w, dist.npart, None))
        dv.free()
    return Spline(final_dist)


def _moments(distribution, vp):
    vp = _f64(vp).ravel()
    dv = DeviceVector(distribution.ctx, vp.size, vp)
    out = np.empty(5)
    check(_lib().vpm_moments(distribution._h, None, dv.ptr, vp.size, _hp(out)))
    dv.free()
    return out


def projection_density(distribution, vp, isDerivative=False):
    m = _moments(distribution, vp)
    return m[3] if isDerivative else m[0]


def projection_momentum(distribution, vp, isDerivative=False):
    m = _moments(distribution, vp)
    return m[4] if isDerivative else m[1]


def projection_energy(distribution, vp, isDerivative=False):
    if isDerivative:
        raise VpmError(-5, "the reference never takes the energy moment of f' (density.jl:15-20)")
    return _moments(distribution, vp)[2]


def compute_f_densities(distribution, vp):
    m = _moments(distribution, vp)
    return m[0], m[1], m[2]


def compute_df_densities(distribution, vp):
    m = _moments(distribution, vp)
    return m[3], m[4]


def compute_coefficients(distribution, particle_dist, vp):
    """lenard_bernstein_conservative.jl:11-21"""
    n, nu, neps, B1, B2 = _moments(distribution, vp)
    B1, B2 = -B1, -B2
    A1 = (neps * B1 - nu * B2) / (n * neps - nu ** 2)
    A2 = -(nu * B1 - n * B2) / (n * neps - nu ** 2)
    return A1, A2


# ------------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------------
class VlasovPoisson:
    def __init__(self, dist, potential):
        self.distribution, self.potential = dist, potential


def update_potential_(model):
    """update_potential!(model): src/models/vlasov_poisson.jl:12-15"""
    d = model.distribution
    check(_lib().vpm_update_potential(model.potential._h, d._h, None, None))


def s_advection_(dist, potential, tau):
    """exact flow of the drift on a device-resident distribution (vlasov_poisson.jl:53-58)"""
    check(_lib().vpm_push_drift(potential._h, dist._h, float(tau)))


def s_acceleration_(dist, potential, tau, field_dist=None, scale=1.0):
    """exact flow of the kick (vlasov_poisson.jl:61-67): field from field_dist (default: dist itself)"""
    src = field_dist or dist
    check(_lib().vpm_update_potential(potential._h, src._h, None, None))
    check(_lib().vpm_push_kick(potential._h, dist._h, None, float(tau), float(scale)))


class Entropy:
    pass


class CollisionEntropy(Entropy):
    """holder of the SplineDistribution (src/entropies/collision_entropy.jl:1-10)"""

    def __init__(self, dist):
        self.dist = dist
        self.cache = {np.float64: dist, float: dist}
        self.value, self.floored = None, None   # set by compute_entropy_ (non-reference diagnostic)


ENTROPY_FLOOR = 1e-14   # f_s below this is round-off of the fp64 projection (f_s = O(0.1) in the bulk)


def compute_entropy_(entropy, dist, velocities=None, f_floor=ENTROPY_FLOOR, project=True):
    """compute_entropy!(entropy, dist) -- a TODO upstream (src/entropies/collision_entropy.jl:12-15), so this is a
    NON-REFERENCE diagnostic: S = -sum_p w_p ln max(f_s(v_p), f_floor), the particle form of -int f ln f dv with f given by
    the spline projection of the particles (project=True re-projects dist first, as update_entropy! does,
    lenard_bernstein.jl:15-17; False uses the spline currently held by entropy.dist).  `velocities`: evaluate at these
    instead of the particles' own.  Sets entropy.value / entropy.floored and returns S."""
    sd = entropy.dist
    _, v, w = dist.ptrs()
    dv = None
    if velocities is not None:
        vel = _f64(velocities).ravel()
        if vel.size != dist.npart:
            raise ValueError("velocities must have one entry per particle")
        dv = DeviceVector(dist.ctx, vel.size, vel)
        v = dv.ptr
    if project:
        check(_lib().vpm_project_v(sd._h, v, w, dist.npart, None))
    S, nf = C.c_double(), C.c_double()
    check(_lib().vpm_entropy_v(sd._h, None, v, w, dist.npart, float(f_floor), C.byref(S), C.byref(nf)))
    if dv is not None:
        dv.free()
    entropy.value, entropy.floored = S.value, int(nf.value)
    return S.value


class LenardBernstein:
    conservative = False

    def __init__(self, dist, ent, nu=1.0):
        self.dist, self.ent, self.nu = dist, ent, float(nu)


class ConservativeLenardBernstein(LenardBernstein):
    conservative = True


def _lb_rhs(vdot, v, params, conservative):
    model = params["model"]
    idist, sdist = params["idist"], model.ent.dist
    v = _f64(v).ravel()
    dv = DeviceVector(idist.ctx, v.size, v)
    out = DeviceVector(idist.ctx, v.size)
    _, _, w = idist.ptrs()
    A = np.zeros(2)
    check(_lib().vpm_lb_rhs(sdist._h, dv.ptr, w, v.size, float(params["nu"]), int(conservative), out.ptr, None, _hp(A)))
    vdot[...] = out.download().reshape(np.shape(vdot))
    dv.free(); out.free()
    return vdot


def LB_rhs_(vdot, v, params, t=0.0):
    """LB_rhs!(v̇, v, params, t): src/models/lenard_bernstein.jl:20-30"""
    return _lb_rhs(vdot, v, params, False)


def CLB_rhs_(vdot, v, params, t=0.0):
    """CLB_rhs!(v̇, v, params, t): src/models/lenard_bernstein_conservative.jl:24-36"""
    return _lb_rhs(vdot, v, params, True)


def LB_rhs_GI_(v, t, q, params):
    """LB_rhs_GI!(v, t, q, params): the GeometricIntegrators argument order (lenard_bernstein.jl:32-34)"""
    return LB_rhs_(v, q, params, t)


def CLB_rhs_GI_(v, t, q, params):
    """CLB_rhs_GI!(v, t, q, params): lenard_bernstein_conservative.jl:38-50"""
    return CLB_rhs_(v, q, params, t)


def LB_rhs(v, params, fs):
    """LB_rhs(v, params, fs::Spline) -> v̇ for a GIVEN spline ("used for plotting", lenard_bernstein.jl:37-44):
    no projection, v is any host grid"""
    v = _f64(v).ravel()
    dfdv = Derivative(1) * fs
    return -float(params["nu"]) * (dfdv(v) + v * fs(v))


def CLB_rhs(v, params, fs):
    """CLB_rhs(v, params, fs::Spline) -> v̇ (lenard_bernstein_conservative.jl:53-64): A from the moments of fs over v"""
    v = _f64(v).ravel()
    dfdv = Derivative(1) * fs
    A1, A2 = compute_coefficients(params["model"].ent.dist, params["idist"], v)
    return -float(params["nu"]) * (dfdv(v) + (A1 + A2 * v) * fs(v))


# ------------------------------------------------------------------------------------------------
# methods (drivers)
# ------------------------------------------------------------------------------------------------
def _ntime(tspan, tstep):
    """number of steps of a tspan: GeometricEquations' ntime = Int(abs(div(tend - tbegin, tstep, RoundUp))) -- a ceiling of
    the floating-point quotient (2.1 / 0.3 = 7.000000000000001 gives 8 steps, as upstream; tspan_for avoids that)"""
    return int(abs(math.ceil((tspan[1] - tspan[0]) / tstep)))


def tspan_for(nsteps, tstep, t0=0.0):
    """a tspan (t0, t1) for which ntime is exactly nsteps (t0 + nsteps * tstep can round to a quotient just above
    nsteps, which the upstream ceiling turns into one more step)"""
    t1 = t0 + nsteps * tstep
    while nsteps > 0 and _ntime((t0, t1), tstep) > nsteps:
        t1 = math.nextafter(t1, t0)
    return (t0, t1)


class H5Writer:
    """The trajectory files of run! (datasets z[nd,np,nt+1] chunk (nd,np,1), splitting.jl:32-34; z[np,nt+1] and t,
    geometric_integrator.jl:21-25) written by the library's own minimal HDF5 writer (csrc/h5min.cpp): h5py / HDF5.jl
    are not needed.  Shapes are given Julia-style (column-major, frame axis last) and reversed for the file exactly
    as HDF5.jl does.  Host-only."""

    def __init__(self, path):
        self._h = C.c_void_p()
        check(_lib().vpm_h5_create(os.fsencode(path), C.byref(self._h)))
        self._ids = {}

    def create_dataset(self, name, julia_shape):
        dims = (C.c_int64 * len(julia_shape))(*reversed([int(d) for d in julia_shape]))
        ident = C.c_int(-1)
        check(_lib().vpm_h5_add_dataset(self._h, name.encode(), len(julia_shape), dims, C.byref(ident)))
        self._ids[name] = ident.value
        return self

    def commit(self):
        check(_lib().vpm_h5_commit(self._h))
        return self

    def write_frame(self, name, frame, data, offset=0):
        a = np.ascontiguousarray(data, dtype=np.float64)
        check(_lib().vpm_h5_write(self._h, self._ids[name], int(frame), int(offset), a.size, _hp(a)))

    def close(self):
        if self._h:
            check(_lib().vpm_h5_close(self._h))
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SplittingMethod:
    """SplittingMethod(model::VlasovPoisson{1,1}, tspan, tstep): Strang splitting (vlasov_poisson.jl:73-89).

    field="frozen" reproduces the shipped behaviour (potential always deposited from model.distribution,
    SURVEY F4); field="selfconsistent" is the physical loop of the legacy integrate_vp! (src/vlasov_poisson.jl:94-115).
    """

    def __init__(self, model, tspan, tstep, field="frozen", chi=1.0):
        if field not in ("frozen", "selfconsistent"):
            raise ValueError("field must be 'frozen' or 'selfconsistent'")
        self.model, self.tspan, self.tstep, self.field, self.chi = model, tspan, float(tstep), field, float(chi)
        self.diagnostics = None


def _vp_steps(method, nsteps, diag_mode):
    d, pot = method.model.distribution, method.model.potential
    diag = np.zeros((nsteps + 1, 3)) if diag_mode else None
    mode = _cabi.VP_FROZEN if method.field == "frozen" else _cabi.VP_SELFCONSISTENT
    check(_lib().vpm_vp_strang_steps(pot._h, d._h, method.tstep, method.chi, int(nsteps), mode, int(diag_mode), _hp(diag)))
    return diag


class GeometricIntegrator:
    """GeometricIntegrator(model::{Conservative}LenardBernstein{1,1}, tspan, tstep): RK438
    (lenard_bernstein.jl:68-84, lenard_bernstein_conservative.jl:88-104)"""

    def __init__(self, model, tspan, tstep):
        self.model, self.tspan, self.tstep = model, tspan, float(tstep)
        self.diagnostics = None


def run_(method, h5file=None, save_stride=None, diag_mode=1, entropy=False, f_floor=ENTROPY_FLOOR):
    """run!(method, h5file): src/methods/splitting.jl:23-52, src/methods/geometric_integrator.jl:12-44.

    The state stays on the device between steps (vpm_vp_run / vpm_lb_run).  With h5file, the frames of steps
    0, k, 2k, ..., nt (k = save_stride, default 1 = the reference's every-step output) are written to an HDF5
    file in the reference's layout (dataset "z", plus "t"), copied off the device while the next steps compute.
    Without h5file no trajectory is kept.  Diagnostics (W,K,M) or (sum v, sum v^2) of every step are kept in
    method.diagnostics; method.frames = number of frames written.  GeometricIntegrator runs with entropy=True also
    record the collision entropy S(t_n) of every step (compute_entropy_: non-reference diagnostic, one extra gather pass
    per step) in method.entropy, and the number of floored particles in method.entropy_floored.
    """
    nt = _ntime(method.tspan, method.tstep)
    stride = int(save_stride or 1) if h5file is not None else 0
    path = os.fsencode(h5file) if h5file is not None else None
    frames = C.c_int(0)
    if isinstance(method, SplittingMethod):
        d, pot = method.model.distribution, method.model.potential
        diag = np.zeros((nt + 1, 3)) if diag_mode else None
        mode = _cabi.VP_FROZEN if method.field == "frozen" else _cabi.VP_SELFCONSISTENT
        check(_lib().vpm_vp_run(pot._h, d._h, method.tstep, method.chi, nt, mode, int(diag_mode), stride, path, _hp(diag),
                                C.byref(frames)))
        method.diagnostics, method.frames = diag, frames.value
        return d
    if isinstance(method, GeometricIntegrator):
        m = method.model
        d, sd = m.dist, m.ent.dist
        diag = np.zeros((nt + 1, 2))
        check(_lib().vpm_vspace_entropy_history(sd._h, int(bool(entropy)), float(f_floor)))
        try:
            check(_lib().vpm_lb_run(sd._h, d._h, m.nu, method.tstep, float(method.tspan[0]), nt, int(m.conservative), stride, path,
                                    _hp(diag), C.byref(frames)))
            method.entropy = method.entropy_floored = None
            if entropy:
                S, nf = np.zeros(nt + 1), np.zeros(nt + 1)
                check(_lib().vpm_vspace_entropy_get(sd._h, _hp(S), _hp(nf), nt + 1))
                method.entropy, method.entropy_floored = S, nf
        finally:
            _lib().vpm_vspace_entropy_history(sd._h, 0, float(f_floor))
        method.diagnostics, method.frames = diag, frames.value
        return d
    raise TypeError("run_ expects a SplittingMethod or a GeometricIntegrator")


# ------------------------------------------------------------------------------------------------
# examples / initial conditions
# ------------------------------------------------------------------------------------------------


class NormalDistribution:
    def __init__(self, domain=(0.0, 1.0)):
        self.domain = domain


class UniformDistribution:
    """UniformDistribution(xdomain, vdomain): x and v uniform (src/examples/uniform.jl:2-8)"""

    def __init__(self, xdomain=(0.0, 1.0), vdomain=(-2.0, 2.0)):
        self.xdomain, self.vdomain = xdomain, vdomain


class ShiftedUniformDistribution:
    """ShiftedUniformDistribution(xdomain, vdomain, shift): v uniform on vdomain + shift (src/examples/shifteduniform.jl:2-9)"""

    def __init__(self, xdomain=(0.0, 1.0), vdomain=(-2.0, 2.0), shift=2.0):
        self.xdomain, self.vdomain, self.shift = xdomain, vdomain, shift


class ShiftedNormalV:
    """ShiftedNormalV(domain, shift): x uniform on domain, v ~ N(shift, 1) (src/examples/shiftednormalv.jl:1-7)"""

    def __init__(self, domain=(-5.0, 5.0), shift=2.0):
        self.domain, self.shift = domain, shift


class BumpOnTail:
    def __init__(self, eps=0.03, kappa=0.3, alpha=0.1, sigma=0.5, v0=4.5):
        self.eps, self.kappa, self.alpha, self.sigma, self.v0 = eps, kappa, alpha, sigma, v0

    @property
    def L(self):
        return 2 * math.pi / self.kappa


class DoubleMaxwellian:
    def __init__(self, domain=(-5.0, 5.0), shift=3.0):
        self.domain, self.shift = domain, shift


def initialize_(dist, params, seed=DEFAULT_SEED, offset=0, ntotal=None, xmax=None):
    """initialize!(dist, example): device-side, counter-based in the global particle index so a slab
    [offset, offset+npart) of an ntotal-particle ensemble is reproducible on any rank.

    NormalDistribution maps x through a data-dependent affine transform (normal.jl:19-21): xmax = ceil(max |x0|)
    is reduced on the device over this rank's particles unless `xmax` is given (multi-rank runs pass one value).
    """
    n = dist.npart
    ntotal = n if ntotal is None else int(ntotal)
    if isinstance(params, BumpOnTail):
        check(_lib().vpm_sample_bump_on_tail(dist._h, int(offset), ntotal, int(seed), params.eps, params.kappa,
                                              params.alpha, params.sigma, params.v0))
    elif isinstance(params, DoubleMaxwellian):
        check(_lib().vpm_sample_maxwellian(dist._h, int(offset), ntotal, int(seed), params.domain[0], params.domain[1],
                                            float(params.shift), 1, 1.0))
    elif isinstance(params, ShiftedNormalV):
        check(_lib().vpm_sample_maxwellian(dist._h, int(offset), ntotal, int(seed), params.domain[0], params.domain[1],
                                            float(params.shift), 0, 1.0))
    elif isinstance(params, (UniformDistribution, ShiftedUniformDistribution)):
        check(_lib().vpm_sample_uniform(dist._h, int(offset), ntotal, int(seed), params.xdomain[0], params.xdomain[1],
                                         params.vdomain[0], params.vdomain[1], float(getattr(params, "shift", 0.0)), 1.0))
    elif isinstance(params, NormalDistribution):
        used = C.c_double()
        check(_lib().vpm_sample_normal(dist._h, int(offset), ntotal, int(seed), params.domain[0], params.domain[1],
                                       float(xmax or 0.0), C.byref(used)))
        dist.xmax = used.value
    else:
        raise TypeError(f"no initialize_ method for {type(params).__name__}")
    return dist
