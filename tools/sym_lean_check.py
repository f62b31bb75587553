"""Slab-free symmetric kernel (vsb_gravity_sym_lean.cu: persistent CTAs, ordered accumulation into
acc[]) against the slab kernel: timing over P-range sizes (VSB_SYM_CHUNK), run-to-run bit identity
(GPU box).
    python tools/sym_lean_check.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_ACC  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [65536, 131072, 1048576]
dev = Device(0)
for n in sizes:
    s = synth.plummer2d(n) if n <= 131072 else synth.rotating_disk(n)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    os.environ["VSB_GRAVITY"] = "rows"
    dev.allpairs_accel(1.0)
    ref = dev.get(F_ACC)
    os.environ["VSB_GRAVITY"] = "sym"
    reps = 20 if n <= 131072 else 3
    for shape in ["slab", "lean"]:
        os.environ["VSB_SYM_MODE"] = shape
        for chunk in ([(0, 0)] if shape == "slab" else [(0, 0), (1, 0), (2, 0), (4, 0)]):
            if chunk[0]:
                os.environ["VSB_SYM_CHUNK"] = str(chunk[0])
            else:
                os.environ.pop("VSB_SYM_CHUNK", None)
            dev.allpairs_accel(1.0)
            acc = dev.get(F_ACC)
            dev.allpairs_accel(1.0)
            same = np.array_equal(acc, dev.get(F_ACC))
            err = np.max(np.linalg.norm(acc - ref, axis=1) / np.linalg.norm(ref, axis=1))
            best = 1e9
            for _ in range(3):
                dev.sync()
                t0 = time.perf_counter()
                for _ in range(reps):
                    dev.allpairs_accel(1.0)
                dev.sync()
                best = min(best, (time.perf_counter() - t0) / reps)
            print("n=%d shape %-9s chunk,- %-8s: %9.3f ms  %.4e pairs/s  pipe %.3f  rel diff vs rows %.1e  "
                  "bit-identical rerun %s" % (n, shape, chunk, best * 1e3, n * (n - 1) / best,
                                              n * (n - 1) * 8 / best / (148 * 64 * 1965e6), err, same), flush=True)
    os.environ.pop("VSB_SYM_MODE", None)
    os.environ.pop("VSB_SYM_CHUNK", None)
