"""ctypes binding of libvpm_b200.so (include/vpm_b200.h).

The library is the product; this module only declares its C ABI.  There is no CPU fallback: if the
shared library is missing or no CUDA device is present the calls fail loudly (VpmError).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VPM_B200_LIB: load another build of the same library (kernel experiments compiled with different macros)
LIB_PATH = os.environ.get("VPM_B200_LIB") or os.path.join(_HERE, "lib", "libvpm_b200.so")

VPM_OK = 0
VP_SELFCONSISTENT = 0
VP_FROZEN = 1

_D = C.POINTER(C.c_double)
_vp = C.c_void_p
_i32, _i64, _u64, _f64 = C.c_int, C.c_int64, C.c_uint64, C.c_double

# name -> (restype, argtypes): exactly the declarations of include/vpm_b200.h
SIGNATURES = {
    "vpm_last_error": (C.c_char_p, []),
    "vpm_version": (_i32, []),
    "vpm_ctx_create": (_i32, [_i32, _vp, C.POINTER(_vp)]),
    "vpm_ctx_destroy": (_i32, [_vp]),
    "vpm_sync": (_i32, [_vp]),
    "vpm_device_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64)]),
    "vpm_launch_count": (_i64, [_vp]),
    "vpm_profile": (_i32, [_vp, _i32]),
    "vpm_profile_get": (_i32, [_vp, _vp, _vp]),
    "vpm_profile_get_lb": (_i32, [_vp, _vp, _vp]),
    "vpm_ctx_bind_numa": (_i32, [_vp, C.c_char_p, _i32]),
    "vpm_host_alloc": (_i32, [_i64, C.POINTER(_vp)]),
    "vpm_host_free": (_i32, [_vp]),
    "vpm_dev_alloc": (_i32, [_vp, _i64, C.POINTER(_vp)]),
    "vpm_dev_free": (_i32, [_vp, _vp]),
    "vpm_memcpy_h2d": (_i32, [_vp, _vp, _vp, _i64]),
    "vpm_memcpy_d2h": (_i32, [_vp, _vp, _vp, _i64]),
    "vpm_particles_create": (_i32, [_vp, _i64, C.POINTER(_vp)]),
    "vpm_particles_destroy": (_i32, [_vp]),
    "vpm_particles_size": (_i64, [_vp]),
    "vpm_particles_ptrs": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "vpm_particles_ptrs_const": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "vpm_particles_upload_aos": (_i32, [_vp, _vp, _i32]),
    "vpm_particles_download_aos": (_i32, [_vp, _vp, _i32]),
    "vpm_particles_upload_soa": (_i32, [_vp, _vp, _vp, _vp]),
    "vpm_particles_download_soa": (_i32, [_vp, _vp, _vp, _vp]),
    "vpm_particles_set_uniform_weight": (_i32, [_vp, _f64]),
    "vpm_sample_bump_on_tail": (_i32, [_vp, _i64, _i64, _u64, _f64, _f64, _f64, _f64, _f64]),
    "vpm_sample_maxwellian": (_i32, [_vp, _i64, _i64, _u64, _f64, _f64, _f64, _i32, _f64]),
    "vpm_resample_v": (_i32, [_vp, _vp, _vp, _i64, _i64, _u64, _i32, _vp]),
    "vpm_sample_uniform": (_i32, [_vp, _i64, _i64, _u64, _f64, _f64, _f64, _f64, _f64, _f64]),
    "vpm_sample_normal": (_i32, [_vp, _i64, _i64, _u64, _f64, _f64, _f64, _D]),
    "vpm_xspace_create": (_i32, [_vp, _f64, _f64, _i32, _i32, C.POINTER(_vp)]),
    "vpm_xspace_destroy": (_i32, [_vp]),
    "vpm_xspace_stencils": (_i32, [_vp, _vp, _vp]),
    "vpm_deposit_x": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "vpm_poisson_solve": (_i32, [_vp, _vp, _vp]),
    "vpm_mass_solve_x": (_i32, [_vp, _vp, _vp]),
    "vpm_gather_x": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "vpm_field_energy": (_i32, [_vp, _vp, _D]),
    "vpm_push_drift": (_i32, [_vp, _vp, _f64]),
    "vpm_push_kick": (_i32, [_vp, _vp, _vp, _f64, _f64]),
    "vpm_update_potential": (_i32, [_vp, _vp, _vp, _vp]),
    "vpm_vp_strang_steps": (_i32, [_vp, _vp, _f64, _f64, _i32, _i32, _i32, _vp]),
    "vpm_vp_strang_steps_async": (_i32, [_vp, _vp, _f64, _f64, _i32, _i32, _i32]),
    "vpm_vp_strang_step_host": (_i32, [_vp, _vp, _vp, _vp, _f64, _f64, _i32]),
    "vpm_xspace_get": (_i32, [_vp, _vp, _vp]),
    "vpm_vspace_create": (_i32, [_vp, _f64, _f64, _i32, _i32, _i32, C.POINTER(_vp)]),
    "vpm_vspace_destroy": (_i32, [_vp]),
    "vpm_vspace_size": (_i32, [_vp]),
    "vpm_vspace_mass": (_i32, [_vp, _vp]),
    "vpm_deposit_v": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "vpm_mass_solve_v": (_i32, [_vp, _vp, _vp]),
    "vpm_project_v": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "vpm_gather_v": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "vpm_moments": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "vpm_lb_rhs": (_i32, [_vp, _vp, _vp, _i64, _f64, _i32, _vp, _vp, _vp]),
    "vpm_lb_rk438_steps": (_i32, [_vp, _vp, _f64, _f64, _i32, _i32, _vp]),
    "vpm_lb_rk438_steps_async": (_i32, [_vp, _vp, _f64, _f64, _i32, _i32]),
    "vpm_vspace_get": (_i32, [_vp, _vp, _vp]),
    "vpm_entropy_v": (_i32, [_vp, _vp, _vp, _vp, _i64, _f64, _D, _D]),
    "vpm_vspace_entropy_history": (_i32, [_vp, _i32, _f64]),
    "vpm_vspace_entropy_get": (_i32, [_vp, _vp, _vp, _i32]),
    "vpm_h5_create": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "vpm_h5_add_dataset": (_i32, [_vp, C.c_char_p, _i32, C.POINTER(_i64), C.POINTER(_i32)]),
    "vpm_h5_commit": (_i32, [_vp]),
    "vpm_h5_write": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp]),
    "vpm_h5_close": (_i32, [_vp]),
    "vpm_vp_run": (_i32, [_vp, _vp, _f64, _f64, _i32, _i32, _i32, _i32, C.c_char_p, _vp, C.POINTER(_i32)]),
    "vpm_lb_run": (_i32, [_vp, _vp, _f64, _f64, _f64, _i32, _i32, _i32, C.c_char_p, _vp, C.POINTER(_i32)]),
    "vpm_galerkin_periodic": (_i32, [_f64, _f64, _i32, _i32, _vp, _vp, _vp]),
    "vpm_galerkin_clamped": (_i32, [_f64, _f64, _i32, _i32, _i32, C.POINTER(_i32), _vp, _vp]),
    "vpm_selftest_wrap": (_i32, [_i32]),
    "vpm_comm_unique_id": (_i32, [_vp]),
    "vpm_comm_init": (_i32, [_vp, _i32, _i32, _vp]),
    "vpm_comm_destroy": (_i32, [_vp]),
    "vpm_p2p_prepare": (_i32, [_vp, _vp]),
    "vpm_p2p_attach": (_i32, [_vp, _i32, _i32, _vp]),
    "vpm_p2p_detach": (_i32, [_vp]),
    "vpm_p2p_error": (_i32, [_vp, C.POINTER(_u64)]),
    "vpm_comm_allreduce": (_i32, [_vp, _vp, _i64]),
}


class VpmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvpm_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libvpm_b200.so; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VpmError(-100, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != VPM_OK:
        msg = lib().vpm_last_error()
        raise VpmError(rc, msg.decode() if msg else "")
    return rc
