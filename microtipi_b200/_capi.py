"""ctypes signatures of include/wfm_b200.h.

``load_library()`` binds the product library ``microtipi_b200/csrc/libwfm_b200.so`` (built by
``__graft_entry__.build()`` with nvcc for sm_100a) and raises if it is missing: there is no CPU
fallback.  Tests may pass another CDLL explicitly (the CPU-emulated build under tests/emu)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwfm_b200.so")

WFM_OK = 0
WFM_ERR_INVALID_ARG = -1
WFM_ERR_UNSUPPORTED = -2
WFM_ERR_STATE = -3
WFM_ERR_CUDA = -4
WFM_ERR_NOMEM = -5
WFM_ERR_INTERNAL = -6

WFM_F64, WFM_F32 = 0, 1
WFM_DEFOCUS, WFM_PHASE, WFM_MODULUS = 0, 1, 2
WFM_J_DEFOCUS, WFM_J_PHASE, WFM_J_MODULUS = 1, 2, 4
WFM_MODULUS_INTENDED, WFM_MODULUS_REFERENCE_LAST_PLANE = 0, 1
WFM_EXCHANGE_HANDLE_BYTES = 64
KERNEL_NAMES = ["psf_pipeline", "jac_pipeline", "jac_reduce", "setters"]

_vp = C.c_void_p
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

SIGNATURES = {
    "wfm_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]),
    "wfm_create_slab": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_double, C.c_int, C.c_int]),
    "wfm_create_batch": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                   C.c_int]),
    "wfm_batch_size": (C.c_int, [_vp]),
    "wfm_batch_set_phase": (C.c_int, [_vp, _vp, C.c_int]),
    "wfm_batch_set_modulus": (C.c_int, [_vp, _vp, C.c_int]),
    "wfm_batch_set_defocus": (C.c_int, [_vp, _vp, C.c_int]),
    "wfm_batch_apply_jacobian": (C.c_int, [_vp, C.c_uint, _vp, _vp]),
    "wfm_create_multi": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _ip, C.c_int]),
    "wfm_multi_parts": (C.c_int, [_vp]),
    "wfm_multi_part_info": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip]),
    "wfm_multi_part": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "wfm_multi_apply_jacobian_dev": (C.c_int, [_vp, C.c_uint, C.POINTER(_vp), _vp]),
    "wfm_exchange_export": (C.c_int, [_vp, C.c_int, _vp]),
    "wfm_exchange_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "wfm_exchange_status": (C.c_int, [_vp]),
    "wfm_exchange_close": (C.c_int, [_vp]),
    "wfm_destroy": (C.c_int, [_vp]),
    "wfm_last_error": (C.c_char_p, [_vp]),
    "wfm_set_stream": (C.c_int, [_vp, _vp]),
    "wfm_synchronize": (C.c_int, [_vp]),
    "wfm_wait_stream": (C.c_int, [_vp, _vp]),
    "wfm_fence_stream": (C.c_int, [_vp, _vp]),
    "wfm_set_optics": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double]),
    "wfm_set_basis": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "wfm_build_basis": (C.c_int, [_vp, C.c_int, C.c_int]),
    "wfm_get_basis": (C.c_int, [_vp, _vp, C.c_int]),
    "wfm_set_phase": (C.c_int, [_vp, _vp, C.c_int]),
    "wfm_set_modulus": (C.c_int, [_vp, _vp, C.c_int]),
    "wfm_set_defocus": (C.c_int, [_vp, _vp, C.c_int]),
    "wfm_set_pupil_arrays": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "wfm_set_modulus_mode": (C.c_int, [_vp, C.c_int]),
    "wfm_get_rho": (C.c_int, [_vp, _vp]),
    "wfm_get_phi": (C.c_int, [_vp, _vp]),
    "wfm_get_psi": (C.c_int, [_vp, _vp]),
    "wfm_get_mask": (C.c_int, [_vp, _vp]),
    "wfm_compute_psf": (C.c_int, [_vp]),
    "wfm_invalidate": (C.c_int, [_vp]),
    "wfm_psf_state": (C.c_int, [_vp]),
    "wfm_get_psf": (C.c_int, [_vp, _vp]),
    "wfm_get_cpx_psf": (C.c_int, [_vp, _vp]),
    "wfm_device_psf": (C.c_int, [_vp, C.POINTER(_vp)]),
    "wfm_device_cpx_psf": (C.c_int, [_vp, C.POINTER(_vp)]),
    "wfm_apply_j_phase": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "wfm_apply_j_defocus": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "wfm_apply_j_modulus": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "wfm_apply_jacobian": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int]),
    "wfm_get_psf_rolled": (C.c_int, [_vp, _vp]),
    "wfm_roll_psf_dev": (C.c_int, [_vp, _vp]),
    "wfm_get_mtf": (C.c_int, [_vp, _vp]),
    "wfm_get_psf_async": (C.c_int, [_vp, _vp]),
    "wfm_wait_transfers": (C.c_int, [_vp]),
    "wfm_apply_j_all": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "wfm_apply_jacobian_dev": (C.c_int, [_vp, C.c_uint, _vp, _vp]),
    "wfm_grad_length": (C.c_int, [_vp]),
    "wfm_conv_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "wfm_conv_create_multi": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_int]),
    "wfm_conv_parts": (C.c_int, [_vp]),
    "wfm_conv_destroy": (C.c_int, [_vp]),
    "wfm_conv_last_error": (C.c_char_p, [_vp]),
    "wfm_conv_set_stream": (C.c_int, [_vp, _vp]),
    "wfm_conv_set_object": (C.c_int, [_vp, _vp]),
    "wfm_conv_set_data": (C.c_int, [_vp, _vp]),
    "wfm_conv_set_weights": (C.c_int, [_vp, _vp]),
    "wfm_conv_cost_and_gradient": (C.c_int, [_vp, C.c_double, _vp, _vp, C.c_int, _dp]),
    "wfm_conv_cost_and_gradient_dev": (C.c_int, [_vp, C.c_double, _vp, _vp, C.c_int, _vp]),
    "wfm_eval_fg": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_double, _dp, _vp]),
    "wfm_fill_uniform": (C.c_int, [_vp, _vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]),
    "wfm_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "wfm_host_free": (C.c_int, [_vp]),
    "wfm_get_info": (C.c_int, [_vp, _ip, _ip, _ip, _ip, _ip, _ip, _ip, _ip, _ip]),
    "wfm_active_extent": (C.c_int, [_vp, _ip, _ip]),
    "wfm_set_profiling": (C.c_int, [_vp, C.c_int]),
    "wfm_get_kernel_times": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "wfm_launch_count": (C.c_uint64, []),
    "wfm_version": (C.c_char_p, []),
}


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach argtypes/restype for every symbol the header declares (raises if one is missing)."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load the sm_100a product library.  Fails loudly when it has not been built."""
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  microtipi_b200 has no CPU fallback.")
    lib = bind(C.CDLL(p))
    if path is None:
        _LIB = lib
    return lib
