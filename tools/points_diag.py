"""N2 diagnostic (GPU box): which vessels of the live-reference fixture vessel_points.npz differ
between the device path and the reference, per result array.  Feeds the exact sets asserted in
tests/test_gpu_kepler.py::test_vessel_points_golden.
    python tools/points_diag.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from volatilespace_b200.device import Device  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "vessel_points.npz"))
import time
dev = Device(0)
nv = len(g["vessel12"])
ecc = g["vessel12"][:, 3]
for exact in (False, True):
  dev.set_march_mode(False, exact_trig=exact)
  res = [np.full((nv, w), np.nan) for w in (2, 2, 2, 6)]
  t0 = time.perf_counter()
  dev.host_vessel_points(g["vessel12"], g["v_ref"], g["body12"], g["body_ref"], *res,
                         steps_limit=int(g["steps_limit"]))
  print("==== exact trigonometry %s: %.3f s, march stats %s" % (exact, time.perf_counter() - t0, dev.march_stats()))
  for name, got in zip(("body_impact", "body_enter_atm", "coi_leave", "coi_enter"), res):
      want = g[name]
      flips = np.flatnonzero((np.isnan(got) != np.isnan(want)).any(axis=1))
      ok = ~np.isnan(want) & ~np.isnan(got)
      rel = np.zeros_like(want)
      rel[ok] = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1.0)
      worst = rel.max(axis=1)
      print("%-15s NaN-pattern flips at vessels %s; rows beyond 1e-11: %s; beyond 2e-7: %s; max rel %.3e" % (
          name, flips.tolist(), np.flatnonzero(worst > 1e-11).tolist(),
          np.flatnonzero(worst > 2e-7).tolist(), worst.max()))
  want, got = g["coi_enter"], res[3]
  body_diff = np.flatnonzero(~((got[:, 0] == want[:, 0]) | (np.isnan(got[:, 0]) & np.isnan(want[:, 0]))))
  print("coi_enter: entered body differs at vessels", body_diff.tolist(), "ecc", np.round(ecc[body_diff], 4).tolist())
  print("high-eccentricity ellipses (0.8 <= e < 1):", np.flatnonzero((ecc >= 0.8) & (ecc < 1)).tolist())
  print("with a predicted entry in the reference:", np.flatnonzero((ecc >= 0.8) & (ecc < 1) & ~np.isnan(want[:, 0])).tolist())
