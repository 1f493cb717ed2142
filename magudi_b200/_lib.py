"""ctypes binding of ``libmagudi_gpu.so`` (C ABI in ``include/magudi_gpu.h``).

The product path has no CPU fallback: loading fails loudly when the shared library has not been
built, and every numerical call fails when no CUDA device is available.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmagudi_gpu.so")


class MagudiGpuError(RuntimeError):
    pass


class Options(C.Structure):
    """``mg_options`` (t_SolverOptions / t_SimulationFlags members read by the hot path)."""
    _fields_ = [
        ("ratioOfSpecificHeats", C.c_double),
        ("viscosityOn", C.c_int),
        ("reynoldsNumberInverse", C.c_double),
        ("prandtlNumberInverse", C.c_double),
        ("powerLawExponent", C.c_double),
        ("bulkViscosityRatio", C.c_double),
        ("dissipationOn", C.c_int),
        ("compositeDissipation", C.c_int),
        ("dissipationAmount", C.c_double),
        ("useTargetState", C.c_int),
        ("useContinuousAdjoint", C.c_int),
        ("steadyStateSimulation", C.c_int),
    ]


_P = C.c_void_p
_I3 = C.POINTER(C.c_int)
_D = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol declared in include/magudi_gpu.h
SIGNATURES = {
    "mg_init": (C.c_int, [C.c_int]),
    "mg_last_error": (C.c_char_p, []),
    "mg_version": (C.c_int, []),
    "mg_synchronize": (C.c_int, []),
    "mg_kernel_launch_count": (C.c_longlong, []),
    "mg_stream_handle": (C.c_void_p, []),
    "mg_profile_enable": (C.c_int, [C.c_int]),
    "mg_profile_get": (C.c_int, [C.c_char_p, _D, C.POINTER(C.c_longlong)]),
    "mg_tuning_set": (C.c_int, [C.c_char_p, C.c_int]),
    "mg_tuning_clear": (C.c_int, []),
    "mg_stencil_create": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "mg_stencil_update": (C.c_int, [_P, C.c_int, _I3, _I3, _I3, C.c_int]),
    "mg_stencil_get_adjoint": (C.c_int, [_P, C.POINTER(_P)]),
    "mg_stencil_destroy": (C.c_int, [_P]),
    "mg_stencil_info": (C.c_int, [_P, _I3]),
    "mg_stencil_coefficients": (C.c_int, [_P, _D, _D, _D, _D]),
    "mg_stencil_apply": (C.c_int, [_P, _P, C.c_int, _I3]),
    "mg_stencil_apply_ghosted": (C.c_int, [_P, _P, C.c_int, _I3, _P, _P]),
    "mg_stencil_apply_interior": (C.c_int, [_P, _P, C.c_int, _I3]),
    "mg_stencil_apply_norm": (C.c_int, [_P, _P, C.c_int, _I3]),
    "mg_stencil_apply_norm_inverse": (C.c_int, [_P, _P, C.c_int, _I3]),
    "mg_stencil_apply_and_project_on_boundary": (C.c_int, [_P, _P, C.c_int, _I3, C.c_int]),
    "mg_stencil_project_on_boundary_and_apply": (C.c_int, [_P, _P, C.c_int, _I3, C.c_int]),
    "mg_grid_create": (C.c_int, [C.c_int, C.c_int, _I3, _I3, _I3, _I3, _D, C.c_int, _I3, _I3, C.POINTER(_P)]),
    "mg_grid_destroy": (C.c_int, [_P]),
    "mg_grid_setup_spatial_discretization": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "mg_grid_set": (C.c_int, [_P, C.c_int, _P]),
    "mg_grid_get": (C.c_int, [_P, C.c_int, _P]),
    "mg_grid_set_iblank": (C.c_int, [_P, _I3]),
    "mg_grid_update": (C.c_int, [_P, _I3]),
    "mg_grid_gradient": (C.c_int, [_P, _P, C.c_int, _P]),
    "mg_grid_inner_product": (C.c_int, [_P, _P, _P, _P, C.c_int, _D]),
    "mg_grid_operator": (_P, [_P, C.c_int, C.c_int]),
    "mg_halo_pack": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "mg_halo_unpack": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "mg_functional_quadrature_on_patches": (C.c_int, [_P, C.c_int, _P, _D]),
    "mg_functional_acoustic_noise": (C.c_int, [_P, C.c_double, _D]),
    "mg_functional_acoustic_noise_forcing": (C.c_int, [_P, C.c_double]),
    "mg_functional_actuator_sensitivity": (C.c_int, [_P, C.c_double, _D]),
    "mg_functional_actuator_gradient": (C.c_int, [_P, C.c_double, _P]),
    "mg_functional_pressure_drag": (C.c_int, [_P, _P, _D]),
    "mg_functional_pressure_drag_forcing": (C.c_int, [_P, _P]),
    "mg_p2p_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "mg_p2p_create_dir": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "mg_p2p_handle_size": (C.c_int, []),
    "mg_p2p_get_handle": (C.c_int, [_P, _P]),
    "mg_p2p_connect": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "mg_p2p_exchange": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "mg_p2p_exchange_overlapped": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "mg_p2p_exchange_masked": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_uint, C.c_int]),
    "mg_p2p_check": (C.c_int, [_P]),
    "mg_p2p_destroy": (C.c_int, [_P]),
    "mg_state_create": (C.c_int, [_P, C.POINTER(Options), C.POINTER(_P)]),
    "mg_state_destroy": (C.c_int, [_P]),
    "mg_state_set": (C.c_int, [_P, C.c_int, _P]),
    "mg_state_get": (C.c_int, [_P, C.c_int, _P]),
    "mg_state_set_async": (C.c_int, [_P, C.c_int, _P]),
    "mg_state_stage_async": (C.c_int, [_P, C.c_int, _P]),
    "mg_state_adopt_staged": (C.c_int, [_P, C.c_int]),
    "mg_state_get_async": (C.c_int, [_P, C.c_int, _P]),
    "mg_state_checkpoint_get_async": (C.c_int, [_P, C.c_int, _P]),
    "mg_transfer_fence": (C.c_int, []),
    "mg_transfer_wait": (C.c_int, []),
    "mg_state_set_time": (C.c_int, [_P, C.c_double]),
    "mg_state_add_acoustic_source": (C.c_int, [_P, _D, C.c_double, C.c_double, C.c_double, C.c_double]),
    "mg_state_update": (C.c_int, [_P]),
    "mg_state_cfl": (C.c_int, [_P, C.c_double, _D]),
    "mg_state_dt": (C.c_int, [_P, C.c_double, _D]),
    "mg_state_checkpoint_store": (C.c_int, [_P, C.c_int]),
    "mg_state_checkpoint_load": (C.c_int, [_P, C.c_int]),
    "mg_state_checkpoint_clear": (C.c_int, [_P]),
    "mg_patch_create": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, _I3, C.c_double, C.c_double, C.POINTER(_P)]),
    "mg_patch_num_points": (C.c_int, [_P, _I3, _I3, _I3]),
    "mg_patch_set_array": (C.c_int, [_P, C.c_char_p, C.c_int, _P]),
    "mg_patch_get_array": (C.c_int, [_P, C.c_char_p, C.c_int, _P]),
    "mg_patch_collect": (C.c_int, [_P, C.c_int, C.c_char_p]),
    "mg_patch_link_interface": (C.c_int, [_P, _P, _P]),
    "mg_functional_accumulate": (C.c_int, [_P, C.c_int, C.c_double, C.c_double]),
    "mg_functional_accumulator_get": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "mg_patch_gradient_buffer_setup": (C.c_int, [_P, C.c_int]),
    "mg_functional_actuator_gradient_record": (C.c_int, [_P, C.c_double, C.POINTER(C.c_int)]),
    "mg_patch_gradient_buffer_flush": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "mg_patch_control_forcing_from_buffer": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "mg_functional_drag_force": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "mg_functional_reynolds_stress": (C.c_int, [_P, _P, _P, C.POINTER(C.c_double)]),
    "mg_functional_reynolds_stress_forcing": (C.c_int, [_P, _P, _P]),
    "mg_functional_momentum_actuator_sensitivity": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "mg_functional_momentum_actuator_gradient": (C.c_int, [_P, C.c_int, _P]),
    "mg_region_set_body_force": (C.c_int, [_P, C.c_int, C.c_double, C.c_double]),
    "mg_region_get_body_force": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mg_rk3_substep": (C.c_int, [_P, C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int, C.c_int]),
    "mg_patch_kolmogorov_setup": (C.c_int, [_P, C.c_double, C.c_int]),
    "mg_patch_set_jet_modes": (C.c_int, [_P, C.c_int, _P]),
    "mg_patch_probe_setup": (C.c_int, [_P, C.c_int]),
    "mg_patch_probe_record": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "mg_patch_probe_flush": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "mg_state_extrema": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), _P, C.POINTER(C.c_double), _P]),
    "mg_state_solution_limit_penalty": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "mg_region_set_solution_limits": (C.c_int, [_P, C.c_int, _P, _P, C.c_double]),
    "mg_region_solution_limit_forcing_switch": (C.c_int, [_P, C.c_int]),
    "mg_state_set_solution_limit_flags": (C.c_int, [_P, C.c_int, C.c_int]),
    "mg_grid_setup_filter": (C.c_int, [_P, C.c_char_p]),
    "mg_state_apply_filter": (C.c_int, [_P, C.c_int, C.c_int]),
    "mg_patch_penalty_amounts": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mg_patch_link_interface_remote": (C.c_int, [_P, _P, C.c_double, C.c_double, C.c_int, C.POINTER(_P)]),
    "mg_region_create": (C.c_int, [C.POINTER(_P)]),
    "mg_region_destroy": (C.c_int, [_P]),
    "mg_region_add_state": (C.c_int, [_P, _P]),
    "mg_region_compute_sponge_strengths": (C.c_int, [_P]),
    "mg_state_sponge_arc_length": (C.c_int, [_P, C.c_int, _P]),
    "mg_state_sponge_strengths_gathered": (C.c_int, [_P, C.c_int, _P]),
    "mg_region_update_patches": (C.c_int, [_P]),
    "mg_region_compute_rhs": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "mg_rk4_substep": (C.c_int, [_P, C.c_int, _D, C.c_double, C.c_int, C.c_int, C.c_int]),
    "mg_rk4_substep_adjoint_phase": (C.c_int, [_P, C.c_int, _D, C.c_double, C.c_int, C.c_int]),
    "mg_region_set_fused": (C.c_int, [_P, C.c_int]),
    "mg_region_uses_fused_rhs": (C.c_int, [_P, C.c_int]),
    "mg_region_uses_fused": (C.c_int, [_P, C.c_int]),
}

_lib = None
_initialised_device = None


def load():
    """Load the shared library (no GPU needed) and attach the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MagudiGpuError(
            f"{LIB_PATH} is missing: build it with `python -m magudi_b200.build` "
            "(libmagudi_gpu has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().mg_last_error()
        raise MagudiGpuError(msg.decode() if msg else f"libmagudi_gpu error {rc}")


def init(device=None):
    """Select the CUDA device of this process (``LOCAL_RANK`` by default)."""
    global _initialised_device
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised_device != device:
        check(lib.mg_init(int(device)))
        _initialised_device = device
    return lib


def lib():
    """The library with a device selected; raises when no CUDA device is available."""
    if _initialised_device is None:
        init()
    return _lib


def i3(v):
    a = (C.c_int * 3)(*[int(x) for x in (list(v) + [1, 1, 1])[:3]])
    return a


def d3(v):
    return (C.c_double * 3)(*[float(x) for x in (list(v) + [0.0, 0.0, 0.0])[:3]])


def fptr(a):
    """Pointer to a Fortran-ordered float64 array (must stay alive during the call)."""
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.ndim == 1)
    return a.ctypes.data_as(C.c_void_p)


def as_f(a, shape=None):
    out = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        out = out.reshape(shape, order="F")
    return out
