"""Game-mode ticks per second at the reference's own scale (the live-reference fixture game_mode:
15 Solar System bodies + 200 vessels): physics_body.move; physics_vessel.move; curve_move of both
(game.py:1248-1296), per call and per tick (GPU box).
    python tools/game_rate.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from volatilespace_b200 import phys_body, phys_vessel  # noqa: E402
from volatilespace_b200._lib import KIND_VESSEL  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402
from volatilespace_b200.synth import calc_orb  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "game_mode.npz"))
dev = Device(0)
P = int(g["P"])
pb = phys_body.Physics(dev, curve_points=P)
pb.load({"gc": float(g["gc"]), "coi_coef": float(g["coi_coef"])}, {"mass": g["mass"]},
        {"a": g["a"], "ecc": g["ecc"], "pe_arg": g["pea"], "ma": g["ma0"], "ref": g["ref"], "dir": g["dr"]})
pb.initial(1)
pv = phys_vessel.Physics(dev, curve_points=P)
pv.load({"gc": float(g["gc"])}, {"mass": g["mass"]}, None, None,
        {"a": g["v_a"], "ecc": g["v_ecc"], "pe_arg": g["v_pea"], "ma": g["v_ma0"], "ref": g["v_ref"],
         "dir": g["v_dr"], "ea": g["v_ea0"]})
orb = calc_orb("vessel", pv.ref, pv.body_mass, None, pv.gc, 0.0, pv.a, pv.ecc)
pv.b, pv.f, pv.n = orb["b"], orb["f"], orb["n"]
dev.kepler_load(KIND_VESSEL, pv.curve_points, pv.a, pv.b, pv.f, pv.ecc, pv.pea, pv.n, pv.dr, pv.ma, pv.ea,
                pv.ref, pv.t)
T = dict.fromkeys(("body.move", "vessel.move", "body.curve_move", "vessel.curve_move"), 0.0)
N = 2000
for it in range(N + 200):
    t0 = time.perf_counter()
    pos, ma, ea = pb.move(1)
    t1 = time.perf_counter()
    pv.move(1, pos)
    t2 = time.perf_counter()
    pb.curve_move()
    t3 = time.perf_counter()
    pv.curve_move()
    t4 = time.perf_counter()
    if it >= 200:
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            T[k] += v
tot = sum(T.values()) / N
print("game tick, %d bodies + %d vessels, P = %d: %.1f us = %.0f ticks/s; per call (us): %s" % (
    len(g["a"]), len(g["v_a"]), P, tot * 1e6, 1 / tot, {k: round(v / N * 1e6, 1) for k, v in T.items()}))
