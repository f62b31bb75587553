"""Batched drop-in for volatilespace/physics/enhanced_kepler_solver.py."""
import numpy as np

from .device import default_device


def solve_kepler_ell(ecc, ma, tol):
    """enhanced_kepler_solver.py:18-69 (ENRKE); scalars or arrays."""
    out = default_device().host_solve_kepler_ell(ecc, ma, tol)
    return float(np.ravel(out)[0]) if np.ndim(ecc) == 0 and np.ndim(ma) == 0 else out
