"""Mid-size ticks of the symmetric slab kernel with and without the tail cut (VSB_SYM_TAIL), CUDA
events per tick, L2 flushed between ticks outside the timed windows (GPU box).
    python tools/midsize_ab.py [n ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import time_ticks_each  # noqa: E402
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_ACC  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402

dev = Device(0)
stream = torch.cuda.ExternalStream(dev.stream_handle())
for n in [int(a) for a in sys.argv[1:]] or [32768, 65536, 98304, 131072, 262144]:
    s = synth.plummer2d(n)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    dev.set_gravity_mode("rows")
    dev.allpairs_accel(1.0)
    ref = dev.get(F_ACC)
    dev.set_gravity_mode("sym")
    for tail in (0, 8):
        os.environ["VSB_SYM_TAIL"] = str(tail)
        dev.upload_bodies(s["mass"], s["pos"], s["vel"])        # the timed ticks below move the bodies
        dev.allpairs_accel(1.0)
        acc = dev.get(F_ACC)
        dev.allpairs_accel(1.0)
        same = np.array_equal(acc, dev.get(F_ACC))
        err = np.max(np.linalg.norm(acc - ref, axis=1) / np.linalg.norm(ref, axis=1))
        dev.upload_bodies(s["mass"], s["pos"], s["vel"])
        for _ in range(3):
            dev.allpairs_tick_async(1.0)
        ms = min(time_ticks_each(torch, stream, lambda: dev.allpairs_tick_async(1.0), 16) / 16 for _ in range(3))
        print("n=%d tail cut %d: %.4f ms per tick, pipe utilisation %.4f, rel diff vs row kernel %.1e, "
              "bit-identical rerun %s" % (n, tail, ms, n * (n - 1) * 8 / (ms * 1e-3) / (148 * 64 * 1965e6), err,
                                          same), flush=True)
os.environ.pop("VSB_SYM_TAIL", None)
