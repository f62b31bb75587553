"""A/B timing of the gravity kernel variants on the 64k Plummer workload (GPU box).
    python tools/gravity_ab.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = Device(0)
print("fp64 peak (TFLOP/s, Gops/s):", [x / 1e12 for x in dev.fp64_peak()])
s = synth.plummer2d(n)
dev.upload_bodies(s["mass"], s["pos"], s["vel"])
ref = None
for tpt in ("1", "2", "4", ""):
    if tpt:
        os.environ["VSB_GRAVITY_TPT"] = tpt
    else:
        os.environ.pop("VSB_GRAVITY_TPT", None)
    dev.allpairs_accel(1.0)
    from volatilespace_b200._lib import F_ACC
    acc = dev.get(F_ACC)
    if ref is None:
        ref = acc
    same = np.array_equal(acc, ref)
    dev.sync()
    reps = 20 if n <= 65536 else 3
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.allpairs_accel(1.0)
    dt = (time.perf_counter() - t0) / reps
    print("TPT=%s  %.3f ms  %.4e pairs/s  identical=%s" % (tpt or "auto", dt * 1e3,
                                                           n * (n - 1) / dt, same))
os.environ["VSB_GRAVITY"] = "sym"
dev.allpairs_accel(1.0)
acc = dev.get(F_ACC)
err = np.max(np.linalg.norm(acc - ref, axis=1) / np.linalg.norm(ref, axis=1))
dev.sync()
t0 = time.perf_counter()
for _ in range(reps):
    dev.allpairs_accel(1.0)
dt = (time.perf_counter() - t0) / reps
print("SYM     %.3f ms  %.4e pairs/s  max row rel diff vs rows kernel %.2e" % (
    dt * 1e3, n * (n - 1) / dt, err))
os.environ.pop("VSB_GRAVITY", None)
print("fp64 peak again:", [x / 1e12 for x in dev.fp64_peak()])
