"""Host-array callback forms and legacy field wrappers of the VP path, over the C ABI.

  vector fields   v_advection!, v_acceleration!, lorentz_force!        src/models/vlasov_poisson.jl:23-50
  flows           s_advection!, s_acceleration!  (z, t, z̄, t̄, params)   src/models/vlasov_poisson.jl:53-67
  field functors  PoissonField, ScaledField, ExternalField, energy     src/electric_field.jl:1-77 (legacy)

These take the integrator state as a host 2 x N matrix exactly like the reference callbacks, so every call
moves the state over PCIe; they exist for drop-in parity.  All per-particle arithmetic still runs on the GPU
(the only host arithmetic is the sign / chi^2 scaling of the n_basis-long coefficient vector).  The
device-resident steppers are in api.py.  `params` is the reference's NamedTuple as a dict:
{"phi": Potential, "model": VlasovPoisson}.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import check
from .api import DeviceVector, ParticleDistribution, _f64, _hp, update_potential_

_vp = C.c_void_p


def _lib():
    return _cabi.lib()


def _gather(potential, x, coefficients, deriv):
    """sum_i c_i B_i^(deriv)(x) for host positions x (one upload, one gather kernel, one download)"""
    ctx = potential.ctx
    x = _f64(x).ravel()
    dx, out = DeviceVector(ctx, x.size, x), DeviceVector(ctx, x.size)
    check(_lib().vpm_gather_x(potential._h, _hp(_f64(coefficients)), dx.ptr, x.size, int(deriv), out.ptr))
    r = out.download()
    dx.free(); out.free()
    return r


def _state_on_device(z, ctx):
    tmp = ParticleDistribution(1, 1, z.shape[1], ctx)
    tmp.upload_aos(np.asarray(z[:2], dtype=float))
    return tmp


def v_advection_(zdot, t, z, params):
    zdot[0, :] = z[1, :]
    zdot[1, :] = 0.0
    return zdot


def v_acceleration_(zdot, t, z, params):
    update_potential_(params["model"])            # deposits from model.distribution
    zdot[0, :] = 0.0
    zdot[1, :] = _gather(params["phi"], z[0, :], -params["phi"].coefficients, 1)
    return zdot


def lorentz_force_(zdot, t, z, params):
    update_potential_(params["model"])
    zdot[0, :] = z[1, :]
    zdot[1, :] = _gather(params["phi"], z[0, :], -params["phi"].coefficients, 1)
    return zdot


def s_advection_host_(z, t, zbar, tbar, params):
    pot = params["phi"]
    tmp = _state_on_device(zbar, pot.ctx)
    check(_lib().vpm_push_drift(pot._h, tmp._h, float(t - tbar)))
    z[:2, :] = tmp.download_aos(2)
    return z


def s_acceleration_host_(z, t, zbar, tbar, params):
    pot = params["phi"]
    update_potential_(params["model"])
    tmp = _state_on_device(zbar, pot.ctx)
    check(_lib().vpm_push_kick(pot._h, tmp._h, None, float(t - tbar), 1.0))
    z[:2, :] = tmp.download_aos(2)
    return z


# ---- legacy field functors (src/electric_field.jl) ------------------------------------------------------
class ElectricField:
    """f(e, x, w, t): update!(f, x, w, t) then efield!(f, e, x)  (electric_field.jl:4-17)"""

    def __call__(self, e, x, w=None, t=0.0):
        if w is not None:
            self.update_(x, w, t)
        return self.efield_(e, x)


class PoissonField(ElectricField):
    """PoissonField(poisson): self-consistent field of the particles (electric_field.jl:39-49)"""

    def __init__(self, potential):
        self.potential = potential
        self._phi = np.zeros(potential.n)

    def update_(self, x, w, t=0.0):
        pot, ctx = self.potential, self.potential.ctx
        x, w = _f64(x).ravel(), _f64(w).ravel()
        dx, dw = DeviceVector(ctx, x.size, x), DeviceVector(ctx, w.size, w)
        check(_lib().vpm_deposit_x(pot._h, dx.ptr, dw.ptr, x.size, None))
        check(_lib().vpm_poisson_solve(pot._h, None, _hp(self._phi)))
        dx.free(); dw.free()

    def efield_(self, e, x, scale=1.0):
        e[...] = _gather(self.potential, x, (-scale) * self._phi, 1).reshape(np.shape(e))
        return e

    def energy(self):
        return self.potential.energy(self._phi)

    def coefficients(self):
        return self._phi


class ExternalField(PoissonField):
    """ExternalField(poisson, coeffs, dt): prescribed time-indexed potential coefficients, column ts =
    round(t/dt) (electric_field.jl:55-69).  coeffs: (n_basis, nt+1)."""

    def __init__(self, potential, coeffs, dt):
        super().__init__(potential)
        self.coeffs, self.dt, self.ts = np.asarray(coeffs, dtype=float), float(dt), 0

    def update_(self, x, w, t=0.0):
        self.ts = int(round(t / self.dt))
        self._phi = np.ascontiguousarray(self.coeffs[:, self.ts])


class ScaledField(ElectricField):
    """ScaledField(field, chi): e ./= chi^2, energy / chi^2 (electric_field.jl:21-35)"""

    def __init__(self, field, chi):
        self.field, self.chi = field, float(chi)

    def update_(self, x, w, t=0.0):
        self.field.update_(x, w, t)

    def efield_(self, e, x):
        return self.field.efield_(e, x, scale=1.0 / self.chi ** 2)

    def energy(self):
        return self.field.energy() / self.chi ** 2

    def coefficients(self):
        return self.field.coefficients()


def ScaledPoissonField(potential, chi):
    return ScaledField(PoissonField(potential), chi)


def ScaledExternalField(potential, coeffs, dt, chi):
    return ScaledField(ExternalField(potential, coeffs, dt), chi)
