"""SHA-256 of the symmetric kernel's accelerations at 1 M and 2.2 M bodies, ptxas object vs re-scheduled
object (GPU box): python tools/sym_bits_large.py"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_ACC  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402
dev = Device(0)
os.environ["VSB_GRAVITY"] = "sym"
for n in (1048576, 2200000):
    s = synth.rotating_disk(n)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    h = []
    for sched in ("ptxas", None):
        if sched: os.environ["VSB_SYM_SCHED"] = sched
        else: os.environ.pop("VSB_SYM_SCHED", None)
        dev.allpairs_accel(1.0)
        h.append(hashlib.sha256(dev.get(F_ACC).tobytes()).hexdigest())
    print(n, "default shape: ptxas", h[0][:16], "re-scheduled", h[1][:16], "IDENTICAL" if h[0] == h[1] else "DIFFERENT", flush=True)
