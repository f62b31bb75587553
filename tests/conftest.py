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
              "add_new_root", "set_mass_new_root", "del_root"]


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
