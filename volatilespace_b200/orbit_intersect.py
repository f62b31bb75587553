"""Batched drop-ins for the numeric routines of volatilespace/physics/orbit_intersect.py and
quartic_solver.py: same names and argument order, arrays of items instead of one
item per call, evaluated by libvsb200.so."""
import numpy as np

from ._lib import check, f64, ptr
from .device import default_device


def solve_quartic(a, b, c, d, e, device=None):
    """quartic_solver.py:57-93 -> complex array (n,4) (or (4,) for scalars)."""
    dev = device or default_device()
    coef = f64(np.stack(np.broadcast_arrays(*[np.atleast_1d(np.asarray(x, dtype=np.float64))
                                              for x in (a, b, c, d, e)]), axis=1))
    roots = np.empty((len(coef), 8))
    check(dev.L.vsb_host_solve_quartic(dev.h, len(coef), ptr(coef), ptr(roots)))
    z = roots[:, 0::2] + 1j * roots[:, 1::2]
    return z[0] if np.ndim(a) == 0 else z


def ell_hyp_intersect_circle(a, b, ecc, c_x, c_y, r, device=None):
    """orbit_intersect.py:17-52.  Scalars -> the reference's variable-length array ([nan] when
    empty); arrays -> (ea (n,4) NaN-padded, count (n,))."""
    dev = device or default_device()
    rows = f64(np.stack(np.broadcast_arrays(*[np.atleast_1d(np.asarray(x, dtype=np.float64))
                                              for x in (a, b, ecc, c_x, c_y, r)]), axis=1))
    ea = np.empty((len(rows), 4))
    cnt = np.empty(len(rows), dtype=np.int64)
    check(dev.L.vsb_host_ell_hyp_intersect_circle(dev.h, len(rows), ptr(rows), ptr(ea), ptr(cnt)))
    if np.ndim(a) == 0:
        return ea[0, :cnt[0]].copy() if cnt[0] else np.array([np.nan])
    return ea, cnt


def predict_enter_coi(vessel_data, body_data, tol, steps_limit, device=None):
    """orbit_intersect.py:125-174.  vessel_data / body_data: one 10-tuple or (n,10) arrays."""
    dev = device or default_device()
    vd = f64(np.atleast_2d(np.asarray(vessel_data, dtype=np.float64)))
    bd = f64(np.atleast_2d(np.asarray(body_data, dtype=np.float64)))
    out = np.empty((len(vd), 5))
    check(dev.L.vsb_host_predict_enter_coi(dev.h, len(vd), ptr(vd), ptr(bd), float(tol),
                                           int(steps_limit), ptr(out)))
    return tuple(out[0]) if np.ndim(vessel_data[0]) == 0 else out
