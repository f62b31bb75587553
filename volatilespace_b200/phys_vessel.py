"""Drop-in for the hot path of volatilespace/physics/phys_vessel.py on a B200:
`Physics.move / curve / curve_move` (phys_vessel.py:477-503, 364-369).  The
COI-crossing prediction, warp checks and curve segmentation of that class are
per-vessel game logic and out of scope.
"""
import numpy as np

from ._lib import KIND_VESSEL, K_EA, K_MA, K_POS
from .device import Device, default_device
from .synth import calc_orb


class Physics:
    """Vessel physics class (game mode), hot path only."""

    def __init__(self, device=None, curve_points=300):
        self.dev = device if isinstance(device, Device) else (
            default_device() if device is None else Device(device))
        self.curve_points = int(curve_points)
        self.t = np.linspace(-np.pi, np.pi, self.curve_points)   # phys_vessel.py:185
        self.gc = 1.0
        self.screen_x, self.screen_y = 1366, 768       # pygame surface size in the reference
        self.n_vessels = 0

    # phys_vessel.py:198-240 (orbit arrays only)
    def load(self, conf, body_data, body_orb, vessel_data, vessel_orb_data):
        self.gc = conf["gc"]
        self.body_mass = np.array(body_data["mass"], dtype=np.float64)
        body_orb = body_orb or {}
        self.body_coi = np.array(body_orb.get("coi", np.zeros(len(self.body_mass))), dtype=np.float64)
        self.names = np.array((vessel_data or {}).get("name", []))
        self.a = np.array(vessel_orb_data["a"], dtype=np.float64)
        self.ecc = np.array(vessel_orb_data["ecc"], dtype=np.float64)
        self.pea = np.array(vessel_orb_data["pe_arg"], dtype=np.float64)
        self.ma = np.array(vessel_orb_data["ma"], dtype=np.float64)
        self.ref = np.array(vessel_orb_data["ref"], dtype=np.int64)
        self.dr = np.array(vessel_orb_data["dir"], dtype=np.float64)
        self.n_vessels = len(self.a)
        self.pos = np.zeros((self.n_vessels, 2))
        self.ea = np.array(vessel_orb_data.get("ea", np.zeros(self.n_vessels)), dtype=np.float64)

    # phys_vessel.py:276-303 (orbit part) -> (vessel_orb, pos, ma, curves_mov)
    def initial(self, warp, body_pos, body_ma=None, body_ea=None):
        orb = calc_orb("vessel", self.ref, self.body_mass, None, self.gc, 0.0, self.a, self.ecc)
        self.b, self.f, self.pe_d, self.ap_d, self.period, self.n, self.u = (
            orb[k] for k in ("b", "f", "pe_d", "ap_d", "period", "n", "u"))
        vessel_orb = {"a": self.a, "ref": np.array(self.ref), "ecc": self.ecc, "pe_d": self.pe_d,
                      "ap_d": self.ap_d, "pea": self.pea, "dir": self.dr, "per": self.period}
        self.dev.kepler_load(KIND_VESSEL, self.curve_points, self.a, self.b, self.f, self.ecc,
                             self.pea, self.n, self.dr, self.ma, self.ea, self.ref, self.t)
        self.move(warp, body_pos, body_ma, body_ea)
        return vessel_orb, self.pos, self.ma, self.curve_move()

    # phys_vessel.py:477-494 -> (pos (Nv,2), ma (Nv,)), refreshed in place
    def move(self, warp, body_pos, body_ma=None, body_ea=None):
        self.body_pos = body_pos
        self.prev_ea = np.array(self.ea)
        self.dev.kepler_move(KIND_VESSEL, warp, body_pos)
        n = self.n_vessels
        self.dev.kepler_get(KIND_VESSEL, K_POS, n, out=self.pos)
        self.dev.kepler_get(KIND_VESSEL, K_MA, n, out=self.ma)
        self.dev.kepler_get(KIND_VESSEL, K_EA, n, out=self.ea)
        return self.pos, self.ma

    # phys_vessel.py:364-369
    def curve(self, vessel):
        return self.dev.host_curve_points(self.ecc[vessel:vessel + 1], self.a[vessel:vessel + 1],
                                          self.b[vessel:vessel + 1], self.pea[vessel:vessel + 1],
                                          self.t)[0]

    # phys_vessel.py:497-503
    def curve_move(self, visible=None):
        self.dev.kepler_curves(KIND_VESSEL)
        if visible is None:
            return self.dev.kepler_curves_get(KIND_VESSEL, 0, self.n_vessels, self.curve_points)
        got = self.dev.kepler_curves_gather(KIND_VESSEL, visible, self.curve_points)
        return {int(i): got[k] for k, i in enumerate(visible)}

    # phys_vessel.py:258-273
    def culling(self, sim_screen, zoom, visible_coi):
        n = self.n_vessels
        dev = self.dev
        self.visible_vessels = dev.kepler_culling(KIND_VESSEL, n, np.full(n, 10.0) / zoom,
                                                  sim_screen, zoom)
        coi_truth = np.zeros(len(self.body_coi), dtype=np.uint8)
        coi_truth[np.asarray(visible_coi, dtype=np.int64)] = 1
        self.visible_orbits = dev.kepler_visible_orbits(KIND_VESSEL, n, coi_truth, self.body_coi,
                                                        self.pe_d, zoom, self.screen_x)
        return self.visible_vessels, self.visible_orbits

    def tick(self, warp, body_pos):
        """move + curve_move in one fused kernel (results stay on the device)."""
        self.dev.kepler_tick(KIND_VESSEL, warp, body_pos)
