"""A/B timing of the symmetric gravity kernel shapes (GPU box).
    python tools/sym_ab.py [n ...]
Shapes are "warps,targets-per-lane,step-unroll,CTAs-per-SM" (VSB_SYM_SHAPE)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_ACC  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402

# the shapes compiled into the library (vsb_gravity_sym.cu, VSB_SYM_CASE list)
SHAPES = os.environ.get("SYM_AB_SHAPES", "8,4,1,2 8,8,1,1 4,8,1,3 4,8,2,3 2,8,1,6").split()
sizes = [int(a) for a in sys.argv[1:]] or [65536, 1048576]
dev = Device(0)
for n in sizes:
    s = synth.plummer2d(n) if n <= 65536 else synth.rotating_disk(n)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    os.environ["VSB_GRAVITY"] = "rows"
    dev.allpairs_accel(1.0)
    ref = dev.get(F_ACC)
    os.environ["VSB_GRAVITY"] = "sym"
    reps = 20 if n <= 65536 else 3
    for shape in SHAPES:
        os.environ["VSB_SYM_SHAPE"] = shape
        try:
            dev.allpairs_accel(1.0)
        except Exception as e:  # noqa: BLE001
            print("n=%d shape %s: %s" % (n, shape, e))
            continue
        acc = dev.get(F_ACC)
        err = np.max(np.linalg.norm(acc - ref, axis=1) / np.linalg.norm(ref, axis=1))
        best = 1e9
        for _ in range(3):
            dev.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                dev.allpairs_accel(1.0)
            dev.sync()
            best = min(best, (time.perf_counter() - t0) / reps)
        print("n=%d shape %-9s %9.3f ms  %.4e pairs/s  max row rel diff vs rows kernel %.2e" % (
            n, shape, best * 1e3, n * (n - 1) / best, err), flush=True)
    os.environ.pop("VSB_SYM_SHAPE", None)
