"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck) on the gravity kernels.
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_ACC  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402

dev = Device(0)
for n in (3000, 20001):
    s = synth.plummer2d(n, seed=4)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    os.environ["VSB_GRAVITY"] = "rows"
    dev.allpairs_accel(1.0)
    ref = dev.get(F_ACC)
    os.environ["VSB_GRAVITY"] = "sym"
    for mode in ("slab", "lean"):
        os.environ["VSB_SYM_MODE"] = mode
        dev.allpairs_accel(1.0)
        acc = dev.get(F_ACC)
        err = np.max(np.linalg.norm(acc - ref, axis=1) / np.linalg.norm(ref, axis=1))
        print("n=%d %s: max row rel diff vs row kernel %.2e" % (n, mode, err), flush=True)
        assert err < 1e-12
    for world in (2, 3):
        os.environ["VSB_SYM_MODE"] = "lean"
        dev.allpairs_sym_partial_async(1.0, 1, world)
        dev.sync()
os.environ.pop("VSB_SYM_MODE", None)
print("done")
