"""Batched drop-ins for the free functions other reference modules import from
volatilespace/physics/phys_shared.py: same names and argument order,
scalar or array arguments, evaluated by libvsb200.so."""
import numpy as np

from .device import default_device


def newton_root_kepler_hyp(ecc, ma, ea_guess):
    """phys_shared.py:114-122."""
    out = default_device().host_newton_root_kepler_hyp(ecc, ma, ea_guess)
    return float(out) if np.ndim(ecc) == 0 and np.ndim(ma) == 0 else out


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
