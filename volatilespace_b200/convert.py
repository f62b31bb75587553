"""Drop-in for volatilespace/physics/convert.py's load/save conversions:
`to_kepler` (convert.py:193-261) and `to_newton` (convert.py:132-190), same signatures and return
dictionaries, evaluated by libvsb200.so.  The reference's O(N^2) pure-Python parent search is
what makes loading maps beyond ~10^4 bodies impractical there."""
import ctypes as C

import numpy as np

from ._lib import check, f64, i64, ptr, VsbError
from .device import default_device


def to_kepler(mass, orb_data, gc, coi_coef, device=None):
    """Newtonian (pos, vel) -> Keplerian elements."""
    dev = device or default_device()
    mass, pos, vel = f64(mass), f64(orb_data["pos"]), f64(orb_data["vel"])
    n = len(mass)
    order = i64(np.argsort(mass)[-1::-1])          # convert.py:207
    ref = np.zeros(n, dtype=np.int64)
    a, ecc, pe, ma, dr = (np.zeros(n) for _ in range(5))
    check(dev.L.vsb_host_to_kepler(dev.h, n, ptr(order), ptr(mass), ptr(pos), ptr(vel), float(gc),
                                   float(coi_coef), ptr(ref), ptr(a), ptr(ecc), ptr(pe), ptr(ma),
                                   ptr(dr)))
    return {"kepler": True, "a": a, "ecc": ecc, "pe_arg": pe, "ma": ma, "ref": ref, "dir": dr}


def to_newton(mass, orb_data, gc, coi_coef, device=None):
    """Keplerian elements -> Newtonian (pos, vel).  Raises ValueError("math domain error") for
    hyperbolic bodies exactly where the reference does (convert.py:170)."""
    dev = device or default_device()
    mass = f64(mass)
    n = len(mass)
    a, ecc, pe, ma, dr = (f64(orb_data[k]) for k in ("a", "ecc", "pe_arg", "ma", "dir"))
    ref = i64(orb_data["ref"])
    pos, vel = np.zeros((n, 2)), np.zeros((n, 2))
    try:
        check(dev.L.vsb_host_to_newton(dev.h, n, ptr(mass), ptr(a), ptr(ecc), ptr(pe), ptr(ma),
                                       ptr(ref), ptr(dr), float(gc), ptr(pos), ptr(vel)))
    except VsbError as e:
        if "math domain error" in str(e):
            raise ValueError("math domain error") from None
        raise
    return {"kepler": False, "pos": pos, "vel": vel}


def _rows(x, width):
    x = np.asarray(x, dtype=np.float64)
    return x.reshape(-1, width) if width > 1 else np.atleast_1d(x)


def kepler_to_velocity(rel_pos, a, ecc, pe_arg, u, dr, device=None):
    """convert.py:20-64: relative velocity at the focus-relative point `rel_pos` of a conic.
    One orbit (2-vector + scalars -> 2-vector) or n of them ((n,2) + (n,) arrays -> (n,2))."""
    dev = device or default_device()
    pos = f64(_rows(rel_pos, 2))
    n = len(pos)
    cols = [f64(np.broadcast_to(_rows(x, 1), (n,))) for x in (a, ecc, pe_arg, u, dr)]
    out = np.empty((n, 2))
    check(dev.L.vsb_host_kepler_to_velocity(dev.h, n, ptr(pos), *(ptr(c) for c in cols), ptr(out)))
    return out[0] if np.ndim(rel_pos) == 1 else out


def velocity_to_kepler(rel_pos, rel_vel, u, failsafe=0, device=None):
    """convert.py:67-131 -> (a, ecc, pe_arg, ma, dr) for one state vector, or an (n,5) array.
    Unlike the reference it leaves the caller's `rel_vel` untouched (the reference rotates it in
    place and only undoes that for failsafe == 2)."""
    dev = device or default_device()
    pos, vel = f64(_rows(rel_pos, 2)), f64(_rows(rel_vel, 2))
    n = len(pos)
    uu = f64(np.broadcast_to(_rows(u, 1), (n,)))
    out = np.empty((n, 5))
    check(dev.L.vsb_host_velocity_to_kepler(dev.h, n, ptr(pos), ptr(vel), ptr(uu), int(failsafe),
                                            ptr(out)))
    if np.ndim(rel_pos) == 1:
        a, ecc, pe, ma, dr = out[0]
        return a, ecc, pe, ma, int(dr)
    return out
