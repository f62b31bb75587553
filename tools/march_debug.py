"""Debug: one long (vessel, body) pair of the vessel_points fixture restarted close to the
disputed COI entry; GPU (parallel and serial) against the C oracle."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
from volatilespace_b200.device import Device
from volatilespace_b200._lib import load, check, ptr
from volatilespace_b200 import orbit_intersect as oi
L = load()

def closed(x0, c, k):
    ks = np.array([k], dtype=np.int64); xs = np.empty(1)
    check(L.vsb_host_running_sum(float(x0), float(c), int(k), 0, 0.0, 1, ptr(ks), ptr(xs), None, None))
    return xs[0]

g = np.load(os.path.join(ROOT, "tests", "golden", "vessel_points.npz"))
v, body, kstart = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
V, B = g["vessel12"][v], g["body12"][body]
a, b, f, ecc, ma, ea, pea, n, dr, pe_d, ap_d, period = V
steps = max(int(max(min(period / B[9] * 5, period / 2), period / 30)), 300)
dt = period / steps
ma1 = closed(ma, dr * n * dt, kstart); bma1 = closed(B[4], B[8] * B[7] * dt, kstart)
ea1 = O.solve_kepler_ell(ecc, ma1); bea1 = O.solve_kepler_ell(B[3], bma1)
for per2 in (3.0e5, 6.0e5):
    vd = np.array([[a, b, f, ecc, ma1, ea1, pea, per2, n, dr]])
    bd = np.array([[B[0], B[1], B[2], B[3], bma1, bea1, B[6], B[7], B[8], B[9]]])
    dev = Device(0)
    par = oi.predict_enter_coi(vd, bd, 1e-5, 300, device=dev)
    dev.set_march_mode(True); ser = oi.predict_enter_coi(vd, bd, 1e-5, 300, device=dev); dev.set_march_mode(False)
    cpu = O.predict_enter_coi(vd[0], bd[0], 1e-5, 300)
    print("period", per2); print(" gpu par", par[0]); print(" gpu ser", ser[0]); print(" cpu    ", np.array(cpu))
# single-step check of the solver on the body's large mean anomaly
m = np.linspace(bma1, bma1 + 40, 2001)
ge = dev.host_solve_kepler_ell(np.full_like(m, B[3]), m, 1e-10)
ce = np.array([O.solve_kepler_ell(B[3], x) for x in m])
print("ENRKE on large ma: max |gpu-cpu|", np.abs(ge - ce).max(), " residual cpu", np.abs(ce - B[3] * np.sin(ce) - m).max(),
      " residual gpu", np.abs(ge - B[3] * np.sin(ge) - m).max())
