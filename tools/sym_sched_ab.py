import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from volatilespace_b200 import synth
from volatilespace_b200.device import Device
dev = Device(0)
os.environ["VSB_GRAVITY"] = "sym"
for n in (65536, 262144, 1048576):
    s = synth.plummer2d(n) if n <= 65536 else synth.rotating_disk(n)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    reps = 20 if n <= 65536 else 5 if n <= 262144 else 3
    for sched in ("ptxas", "resched", "ptxas", "resched"):
        if sched == "ptxas":
            os.environ["VSB_SYM_SCHED"] = "ptxas"
        else:
            os.environ.pop("VSB_SYM_SCHED", None)
        dev.allpairs_accel(1.0)
        best = 1e9
        for _ in range(3):
            dev.sync(); t0 = time.perf_counter()
            for _ in range(reps): dev.allpairs_accel(1.0)
            dev.sync(); best = min(best, (time.perf_counter() - t0) / reps)
        print("n=%d %-8s %9.3f ms  %.4e pairs/s" % (n, sched, best * 1e3, n * (n - 1) / best), flush=True)
