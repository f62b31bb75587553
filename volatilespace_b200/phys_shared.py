"""Batched drop-ins for the free functions other reference modules import from
volatilespace/physics/phys_shared.py: same names and argument order,
scalar or array arguments, evaluated by libvsb200.so."""
import numpy as np

from .device import default_device


def newton_root_kepler_hyp(ecc, ma, ea_guess):
    """phys_shared.py:114-122."""
    out = default_device().host_newton_root_kepler_hyp(ecc, ma, ea_guess)
    return float(np.ravel(out)[0]) if np.ndim(ecc) == 0 and np.ndim(ma) == 0 else out


def orb2xy(a, b, f, ecc, pea, ref_pos, ea):
    """phys_shared.py:187-199: position on the orbit at eccentric anomaly `ea`, offset by the
    reference body's position ((0, 0) when ea == 0 exactly, as there).  Scalars + one 2-vector,
    or arrays + (n,2)."""
    from ._lib import check, f64, ptr
    dev = default_device()
    rp = f64(np.asarray(ref_pos, dtype=np.float64).reshape(-1, 2))
    n = len(rp)
    cols = [f64(np.broadcast_to(np.atleast_1d(np.asarray(x, dtype=np.float64)), (n,)))
            for x in (a, b, f, ecc, pea)]
    eas = f64(np.broadcast_to(np.atleast_1d(np.asarray(ea, dtype=np.float64)), (n,)))
    out = np.empty((n, 2))
    check(dev.L.vsb_host_orb2xy(dev.h, n, *(ptr(c) for c in cols), ptr(rp), ptr(eas), ptr(out)))
    return out[0] if np.ndim(a) == 0 else out


def curve_points(ecc, a, b, pea, t):
    """phys_shared.py:202-216: (P,2) for scalar elements, (N,P,2) for arrays."""
    out = default_device().host_curve_points(ecc, a, b, pea, t)
    return out[0] if np.ndim(a) == 0 else out


def curve_move_to(curves, body_pos, ref, f, pea):
    """phys_shared.py:219-222."""
    return default_device().host_curve_move_to(curves, body_pos, ref, f, pea)


def culling(coords, radius, screen_bounds, zoom):
    """phys_shared.py:224-231: boolean mask of the objects that are on screen."""
    return default_device().host_culling(coords, radius, screen_bounds, zoom)
