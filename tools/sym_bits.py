"""SHA-256 of the symmetric kernel's accelerations for every compiled shape at a few sizes (GPU box).
Used to show that a re-scheduled library (tools/sass_resched.py) computes the same bits:
    python tools/sym_bits.py > a.txt ; <swap the library> ; python tools/sym_bits.py > b.txt ; cmp a.txt b.txt"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_ACC  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402

SHAPES = "8,4,1,2 8,8,1,1 4,8,1,3 4,8,2,3 2,8,1,6 1,8,1,12".split()
dev = Device(0)
os.environ["VSB_GRAVITY"] = "sym"
for n in (8192, 20000, 65536, 262144):
    s = synth.plummer2d(n, seed=n) if n <= 65536 else synth.rotating_disk(n)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    for shape in SHAPES:
        os.environ["VSB_SYM_SHAPE"] = shape
        dev.allpairs_accel(1.0)
        print(n, shape, hashlib.sha256(dev.get(F_ACC).tobytes()).hexdigest(), flush=True)
