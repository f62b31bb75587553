"""A/B of the tail split of the symmetric slab kernel (VSB_SYM_TAIL pieces for the tasks of the last
two waves), one GPU standing in for rank 0 of `world` (GPU box).
    python tools/tail_ab.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_ACC  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402

dev = Device(0)
for n in (1048576, 262144):
    s = synth.rotating_disk(n)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    dev.set_gravity_mode("sym")
    for world in (1, 8):
        for tail in (0, 4, 8, 16):
            os.environ["VSB_SYM_TAIL"] = str(tail)
            best = 1e9
            for rep in range(3):
                dev.sync()
                t0 = time.perf_counter()
                for _ in range(2 if world == 1 else 6):
                    if world == 1:
                        dev.allpairs_accel(1.0)
                    else:
                        dev.allpairs_sym_partial_async(1.0, 0, world)
                dev.sync()
                best = min(best, (time.perf_counter() - t0) / (2 if world == 1 else 6))
            print("n=%d rank 0 of %d, tail split %2d: %9.3f ms (ideal share of the 1-GPU time: x%.4f)" % (
                n, world, tail, best * 1e3, 1.0 / world), flush=True)
