"""ctypes binding of libvsb200.so -- the reference-side stub of INTEGRATION.md.

One entry per symbol declared in include/vsb200.h.  Loading fails loudly when
the CUDA library is missing or cannot be loaded: there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvsb200.so")

VSB_OK, VSB_ERR_CUDA, VSB_ERR_ARG, VSB_ERR_STATE, VSB_ERR_NOMEM = 0, -1, -2, -3, -4

# enum vsb_body_field
F_MASS, F_DEN, F_RAD, F_COI, F_POS, F_VEL, F_REL_VEL, F_ACC, F_PARENTS, F_FOCUS, F_ECC_V, \
    F_SEMI_MAJOR, F_SEMI_MINOR, F_PERIAPSIS_ARG = range(14)
BODY_FIELD_WIDTH = {F_MASS: 1, F_DEN: 1, F_RAD: 1, F_COI: 1, F_POS: 2, F_VEL: 2, F_REL_VEL: 2,
                    F_ACC: 2, F_PARENTS: 1, F_FOCUS: 1, F_ECC_V: 2, F_SEMI_MAJOR: 1,
                    F_SEMI_MINOR: 1, F_PERIAPSIS_ARG: 1}
# enum vsb_kepler_field
K_A, K_B, K_F, K_ECC, K_PEA, K_N, K_DR, K_MA, K_EA, K_POS, K_REF = range(11)
KIND_BODY, KIND_VESSEL = 0, 1


class VsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libvsb200 error %d: %s" % (code, msg))
        self.code = code


_ctx = C.c_void_p
_i64 = C.c_int64
_f64 = C.c_double
_pf = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)
_pv = C.c_void_p

# name -> argtypes (restype is int unless stated)
SIGNATURES = {
    "vsb_abi_version": [],
    "vsb_device_count": [C.POINTER(C.c_int)],
    "vsb_ctx_create": [C.c_int, C.POINTER(_ctx)],
    "vsb_ctx_destroy": [_ctx],
    "vsb_ctx_sync": [_ctx],
    "vsb_ctx_stream": [_ctx, C.POINTER(C.c_void_p)],
    "vsb_ctx_kernel_launches": [_ctx, _pi],
    "vsb_measure_fp64_peak": [_ctx, _pf, _pf],
    "vsb_host_register": [_pv, _i64],
    "vsb_host_unregister": [_pv],
    "vsb_bodies_resize": [_ctx, _i64],
    "vsb_bodies_count": [_ctx, _pi],
    "vsb_bodies_set": [_ctx, C.c_int, _pv, _i64, _i64],
    "vsb_bodies_get": [_ctx, C.c_int, _pv, _i64, _i64],
    "vsb_bodies_device_ptr": [_ctx, C.c_int, C.POINTER(C.c_void_p)],
    "vsb_bodies_set_rows": [_ctx, _i64, _i64],
    "vsb_find_parents": [_ctx, _pv, _pv],
    "vsb_gravity_ref": [_ctx, _f64, _pv],
    "vsb_gravity_ref_begin_async": [_ctx, _pv, C.c_int, C.c_int],
    "vsb_find_parents_best_ptr": [_ctx, C.POINTER(C.c_void_p)],
    "vsb_gravity_ref_finish": [_ctx, _f64],
    "vsb_allpairs_accel": [_ctx, _f64],
    "vsb_allpairs_integrate": [_ctx],
    "vsb_allpairs_step": [_ctx, _f64, _i64],
    "vsb_allpairs_tick_async": [_ctx, _f64],
    "vsb_allpairs_accel_async": [_ctx, _f64],
    "vsb_set_gravity_mode": [_ctx, C.c_int],
    "vsb_allpairs_sym_partial_async": [_ctx, _f64, C.c_int, C.c_int],
    "vsb_allpairs_integrate_async": [_ctx],
    "vsb_editor_ticks": [_ctx, _f64, _f64, _pv, _i64, C.c_int, _pi, _pi, _pi, _pv],
    "vsb_bodies_get_many": [_ctx, C.POINTER(C.c_int), C.c_int, _pv],
    "vsb_kepler_basic": [_ctx, _f64, _f64],
    "vsb_simplified_orbit_coi": [_ctx, _f64, _f64, _pv, _pi],
    "vsb_check_collision": [_ctx, _pi, _pi],
    "vsb_check_collision_part_async": [_ctx, C.c_int, C.c_int],
    "vsb_collision_key_ptr": [_ctx, C.POINTER(C.c_void_p)],
    "vsb_check_collision_finish": [_ctx, _pi, _pi],
    "vsb_delete_row": [_ctx, _i64],
    "vsb_swap_rows": [_ctx, _i64, _i64],
    "vsb_translate_to": [_ctx, _i64],
    "vsb_editor_curve": [_ctx, _i64, _pv, _pv, _pv],
    "vsb_host_allpairs_step": [_ctx, _i64, _pv, _pv, _pv, _f64, _i64],
    "vsb_host_allpairs_accel": [_ctx, _i64, _pv, _pv, _f64, _pv],
    "vsb_host_find_parents": [_ctx, _i64, _pv, _pv, _pv, _pv],
    "vsb_host_check_collision": [_ctx, _i64, _pv, _pv, _pv, _pi, _pi],
    "vsb_host_solve_kepler_ell": [_ctx, _i64, _pv, _pv, _f64, _pv],
    "vsb_host_newton_root_kepler_hyp": [_ctx, _i64, _pv, _pv, _pv, _pv],
    "vsb_host_curve_points": [_ctx, _i64, _i64, _pv, _pv, _pv, _pv, _pv, _pv],
    "vsb_host_curve_move_to": [_ctx, _i64, _i64, _i64, _pv, _pv, _pv, _pv, _pv, _pv],
    "vsb_kepler_load": [_ctx, C.c_int, _i64, _i64] + [_pv] * 11,
    "vsb_kepler_set_order": [_ctx, _pv],
    "vsb_kepler_set": [_ctx, C.c_int, C.c_int, _pv, _i64, _i64],
    "vsb_kepler_get": [_ctx, C.c_int, C.c_int, _pv, _i64, _i64],
    "vsb_kepler_device_ptr": [_ctx, C.c_int, C.c_int, C.POINTER(C.c_void_p)],
    "vsb_kepler_move": [_ctx, C.c_int, _f64, _pv, _i64],
    "vsb_kepler_move_state": [_ctx, C.c_int, _f64, _pv, _i64, _pv, _pv, _pv],
    "vsb_kepler_curves": [_ctx, C.c_int],
    "vsb_kepler_tick": [_ctx, C.c_int, _f64, _pv, _i64],
    "vsb_kepler_tick_async": [_ctx, C.c_int, _f64],
    "vsb_kepler_free": [_ctx, C.c_int],
    "vsb_kepler_curves_get": [_ctx, C.c_int, _i64, _i64, _pv],
    "vsb_kepler_curve_move": [_ctx, C.c_int, _pv],
    "vsb_kepler_curves_device_ptr": [_ctx, C.c_int, C.POINTER(C.c_void_p)],
    "vsb_host_to_kepler": [_ctx, _i64, _pv, _pv, _pv, _pv, _f64, _f64] + [_pv] * 6,
    "vsb_host_to_newton": [_ctx, _i64] + [_pv] * 7 + [_f64, _pv, _pv],
    "vsb_host_solve_quartic": [_ctx, _i64, _pv, _pv],
    "vsb_host_ell_hyp_intersect_circle": [_ctx, _i64, _pv, _pv, _pv],
    "vsb_host_predict_enter_coi": [_ctx, _i64, _pv, _pv, _f64, _i64, _pv],
    "vsb_host_vessel_points": [_ctx, _i64, _i64, _pv, _pv, _pv, _pv, _i64, _pv, _i64, _i64, _i64,
                               C.c_int, _i64, _pv, _pv, _pv, _pv],
    "vsb_host_check_warp": [_ctx, _i64] + [_pv] * 9 + [_f64, _pv, _pv],
    "vsb_host_enter_coi_service": [_ctx, _i64, _pv, _pv, _pv, _pv, C.c_int, _pv],
    "vsb_host_cross_scan": [_ctx, _i64] + [_pv] * 7 + [_pv],
    "vsb_host_cross_apply": [_ctx, _i64, _pv, _pv, _pv, _pv, _i64, _pv, _f64, _pv, _pv],
    "vsb_vessel_events_load": [_ctx] + [_pv] * 9,
    "vsb_vessel_events_set_orbit": [_ctx, _i64, _f64, _f64, _f64],
    "vsb_vessel_events_get": [_ctx, _i64, _i64] + [_pv] * 6,
    "vsb_vessel_points": [_ctx, _i64, _pv, _pv, _i64, _pv, _i64, _i64, _i64, C.c_int, _i64],
    "vsb_vessel_check_warp": [_ctx, _f64, _pv],
    "vsb_vessel_enter_coi_service": [_ctx, C.c_int, _pv, _pi],
    "vsb_vessel_cross_scan": [_ctx, _pv],
    "vsb_host_kepler_to_velocity": [_ctx, _i64] + [_pv] * 7,
    "vsb_host_velocity_to_kepler": [_ctx, _i64, _pv, _pv, _pv, C.c_int, _pv],
    "vsb_host_orb2xy": [_ctx, _i64] + [_pv] * 8,
    "vsb_host_kepler_basic": [_ctx, _i64, _pv, _pv, _pv, _pv, _f64, _f64] + [_pv] * 6,
    "vsb_set_march_mode": [_ctx, C.c_int],
    "vsb_host_running_sum": [_f64, _f64, _i64, C.c_int, _f64, _i64, _pv, _pv, _pv, _pv],
    "vsb_march_stats": [_ctx, _pv],
    "vsb_host_body_thermal": [_ctx, _i64] + [_pv] * 16,
    "vsb_host_culling": [_ctx, _i64, _pv, _pv, _pv, _f64, _pv],
    "vsb_kepler_culling": [_ctx, C.c_int, _pv, _pv, _f64, _pv, _pi],
    "vsb_kepler_culling_uniform": [_ctx, C.c_int, _f64, _pv, _f64, _pv, _pi],
    "vsb_kepler_visible_orbits": [_ctx, C.c_int, _i64, _pv, _pv, _pv, _f64, _f64, _pv, _pi],
    "vsb_kepler_curves_gather": [_ctx, C.c_int, _pv, _i64, _pv],
    "vsb_vessel_curve_segments": [_ctx, _pv, _i64, _pv, _pv, _pv, _pv, _pv, _pv],
}

_lib = None


def load():
    """Load libvsb200.so (raises if it has not been built: run
    `python -m volatilespace_b200.build` or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m volatilespace_b200.build` "
                          "(no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError if the header and the library disagree
        fn.argtypes = args
        fn.restype = C.c_int
    L.vsb_last_error.argtypes = []
    L.vsb_last_error.restype = C.c_char_p
    L.vsb_sym_schedule.argtypes = []
    L.vsb_sym_schedule.restype = C.c_char_p
    _lib = L
    return L


def sym_schedule():
    """"resched" or "ptxas (<why>)": the schedule of the symmetric pair kernel in this library."""
    return load().vsb_sym_schedule().decode("utf-8", "replace")


def check(code):
    if code != VSB_OK:
        raise VsbError(code, load().vsb_last_error().decode("utf-8", "replace"))


def f64(a):
    """C-contiguous float64 view/copy of a (what every entry point expects)."""
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
