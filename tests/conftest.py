import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "live: needs /root/reference (dev container only)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.lib()
    return orc

EDIT_STEPS = ["load", "add_light", "set_mass", "set_pos", "set_vel", "set_den", "del_ordinary",
              "add_new_root", "set_mass_new_root", "del_root", "kepler_inverse_ma",
              "kepler_inverse_ta_apd"]


def new_body(m, p, v):
    return {"name": "new", "mass": m, "density": 3.0e7, "position": np.array(p, float),
            "velocity": np.array(v, float)}


def apply_edit(ph, k):
    """Step k of oracle/gen_golden.py:gen_editor_edits."""
    if k == 1:
        ph.add_body(new_body(3.0, [9000.0, 2000.0], [0.1, 1.2]))
    elif k == 2:
        ph.set_body_mass(7, 1500.0)
    elif k == 3:
        ph.set_body_pos(9, np.array([-4000.0, 7000.0]))
    elif k == 4:
        ph.set_body_vel(9, np.array([0.9, 0.4]))
    elif k == 5:
        ph.set_body_den(3, 0.01)
    elif k == 6:
        ph.del_body(5)
    elif k == 7:
        ph.add_body(new_body(9.0e4, [15000.0, -8000.0], [0.0, 0.3]))
    elif k == 8:
        ph.set_body_mass(4, 5.0e5)
    elif k == 9:
        ph.del_body(0)
    elif k == 10:      # orbit typed into the editor panel: e, omega, pe_d, M (deg), no ta / ap_d
        ph.kepler_inverse(6, 0.3, 40.0, 2000.0, 75.0, 0.0, 0.0, 1.0)
    elif k == 11:      # true anomaly and apoapsis given, clockwise
        ph.kepler_inverse(8, 0.5, 200.0, 1500.0, 10.0, 120.0, 5200.0, -1.0)


# ---- vessel_events fixture helpers (column layout written by oracle/gen_golden.py) ----------
# state columns: a b f ecc pea n dr ma ea prev_ea u pe_d ap_d period
VS = dict(a=0, b=1, f=2, ecc=3, pea=4, n=5, dr=6, ma=7, ea=8, prev_ea=9, u=10, pe_d=11, ap_d=12,
          period=13)


def split_points(p):
    """(nv,12) -> body_impact (nv,2), body_enter_atm (nv,2), coi_leave (nv,2), coi_enter (nv,6)."""
    import numpy as np
    return tuple(np.ascontiguousarray(x) for x in (p[:, 0:2], p[:, 2:4], p[:, 4:6], p[:, 6:12]))


def events_body7(g, bpos):
    """Body rows of the cross_coi orbit change: a ecc pea dr mass pos_x pos_y."""
    import numpy as np
    bs = g["body_static"]
    return np.stack([bs[:, 0], bs[:, 3], bs[:, 4], bs[:, 6], g["body_mass"], bpos[:, 0], bpos[:, 1]],
                    axis=1)


def events_body12(g, bma, bea):
    """Body rows of points(): a b f ecc ma ea pea n dr coi size atm."""
    import numpy as np
    bs = g["body_static"]
    return np.stack([bs[:, 0], bs[:, 1], bs[:, 2], bs[:, 3], bma, bea, bs[:, 4], bs[:, 5], bs[:, 6],
                     bs[:, 7], bs[:, 8], bs[:, 9]], axis=1)


def events_vessel12(v):
    """State columns -> the vessel rows of points(): a b f ecc ma ea pea n dr pe_d ap_d period."""
    import numpy as np
    return np.stack([v[:, 0], v[:, 1], v[:, 2], v[:, 3], v[:, 7], v[:, 8], v[:, 4], v[:, 5], v[:, 6],
                     v[:, 11], v[:, 12], v[:, 13]], axis=1)


def bits_xor(a):
    """Order-independent fingerprint of an array's exact bits (oracle/gen_golden.py:bits_xor)."""
    import numpy as np
    return np.bitwise_xor.reduce(np.ascontiguousarray(a, dtype=np.float64).reshape(-1).view(np.uint64))
