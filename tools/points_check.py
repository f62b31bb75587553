"""Diagnostics for the N2 orbit-points path: error distribution against the golden fixture,
march statistics and timings (GPU box)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200.device import Device

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "vessel_points.npz"))
dev = Device(0)
nv = len(g["vessel12"])
nan = lambda w: np.full((nv, w), np.nan)
bi, ba, cl, ce = nan(2), nan(2), nan(2), nan(6)
for rep in range(2):
    t = time.perf_counter()
    dev.host_vessel_points(g["vessel12"], g["v_ref"], g["body12"], g["body_ref"], bi, ba, cl, ce,
                           steps_limit=int(g["steps_limit"]))
    print("points(all %d vessels): %.3f s  stats %s" % (nv, time.perf_counter() - t, dev.march_stats()))
for got, key in ((bi, "body_impact"), (ba, "body_enter_atm"), (cl, "coi_leave"), (ce, "coi_enter")):
    want = g[key]
    mis = np.flatnonzero(np.isnan(got[:, 0]) != np.isnan(want[:, 0]))
    print(key, "found/not-found differs on vessels", mis.tolist(), " ecc", g["vessel12"][mis, 3].round(3).tolist())
    if key == "coi_enter":
        both = ~np.isnan(got[:, 0]) & ~np.isnan(want[:, 0])
        print("   entered body differs on", np.flatnonzero(both & (got[:, 0] != want[:, 0])).tolist())
        bad = np.flatnonzero(both & (np.abs(got[:, 5] - want[:, 5]) > 1e-6 * np.abs(want[:, 5])))
        for v in bad: print("   v", v, "got", got[v], "want", want[v])
    ok = ~np.isnan(want) & ~np.isnan(got)
    rel = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-300)
    if rel.size:
        print("   n=%d  max rel %.3e  median %.3e  frac<1e-11 %.3f  frac<1e-9 %.3f  exact %.3f" % (
            rel.size, rel.max(), np.median(rel), (rel < 1e-11).mean(), (rel < 1e-9).mean(),
            (rel == 0).mean()))
if "--seq" in sys.argv:
    dev.set_march_mode(True)
    t = time.perf_counter()
    cheap = [v for v in range(nv) if v not in (2, 5, 6, 16)]      # minutes each in the serial mode
    b2, a2, l2, e2 = nan(2), nan(2), nan(2), nan(6)
    dev.host_vessel_points(g["vessel12"], g["v_ref"], g["body12"], g["body_ref"], b2, a2, l2, e2,
                           sel=cheap, steps_limit=int(g["steps_limit"]))
    print("serial mode (%d vessels): %.3f s" % (len(cheap), time.perf_counter() - t))
    dev.set_march_mode(False)
    same = np.array_equal(e2[cheap].view(np.uint64), ce[cheap].view(np.uint64))
    print("parallel == serial bits:", same)
