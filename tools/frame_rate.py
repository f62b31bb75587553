"""Editor frames per second on the built-in Solar System map (config 1: 15 bodies, MODE_REF), the fused
driver with the frame record written into mapped host memory (default) and through a copy call
(VSB_NO_MAPPED_FRAME=1), and the per-method path (GPU box).
    python tools/frame_rate.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from volatilespace_b200.device import Device  # noqa: E402
from volatilespace_b200.phys_editor import Physics  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "solar_editor.npz"))
dev = Device(0)


def rate(name, frames, env=None):
    if env:
        os.environ[env] = "1"
    ph = Physics(dev, curve_points=64)
    ph.load_system({"gc": float(g["gc"]), "rad_mult": float(g["rad_mult"]), "coi_coef": float(g["coi_coef"])},
                   {"mass": g["mass"], "den": g["den"]}, {"pos": g["pos0"], "vel": g["vel0"]})
    ph.kepler_basic()
    fn = getattr(ph, name)
    for _ in range(200):
        fn()
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(frames):
            fn()
        best = max(best, frames / (time.perf_counter() - t0))
    if env:
        os.environ.pop(env)
    return best, ph.pos.copy()


a, pa = rate("frame", 5000)
b, pb = rate("frame", 5000, "VSB_NO_MAPPED_FRAME")
c, _ = rate("frame_calls", 500)
print("fused frame, mapped host record: %.0f frames/s; through a copy call: %.0f frames/s; per-method calls: "
      "%.0f frames/s; identical positions after the same number of frames: %s" % (a, b, c, np.array_equal(pa, pb)))

# the editor also asks for the orbit curves of every body on every frame (editor.py:1604)
ph = Physics(dev, curve_points=300)
ph.load_system({"gc": float(g["gc"]), "rad_mult": float(g["rad_mult"]), "coi_coef": float(g["coi_coef"])},
               {"mass": g["mass"], "den": g["den"]}, {"pos": g["pos0"], "vel": g["vel0"]})
ph.kepler_basic()
for _ in range(100):
    ph.frame()
    ph.curve()
t0 = time.perf_counter()
for _ in range(3000):
    ph.curve()
t1 = time.perf_counter()
for _ in range(3000):
    ph.frame()
    ph.curve()
t2 = time.perf_counter()
print("curve() of 15 bodies x 300 points: %.1f us per call; frame() + curve(): %.1f us = %.0f frames/s" % (
    (t1 - t0) / 3000 * 1e6, (t2 - t1) / 3000 * 1e6, 3000 / (t2 - t1)))
