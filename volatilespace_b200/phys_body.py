"""Drop-in for the hot path of volatilespace/physics/phys_body.py on a B200.

`Physics.load / initial / move / curve / curve_move` keep the reference's names,
argument meaning and return shapes (phys_body.py:177-198, 221-346, 405-411); the
Kepler solve, position compose and curve generation run in libvsb200.so.  The
body-data part of `initial()` (radius, thermal, colour, classification --
phys_body.py:227-282) is UI data and stays with the caller.
"""
import numpy as np

from . import _lib
from ._lib import KIND_BODY, K_EA, K_MA, K_POS
from .device import Device, default_device
from .synth import calc_orb


class Physics:
    """Body physics class (game mode)."""

    def __init__(self, device=None, curve_points=300):
        self.dev = device if isinstance(device, Device) else (
            default_device() if device is None else Device(device))
        self.curve_points = int(curve_points)
        self.t = np.linspace(-np.pi, np.pi, self.curve_points)   # phys_body.py:165
        self.gc, self.coi_coef = 1.0, 0.7
        self.n_bodies = 0

    def reload_settings(self, curve_points=None):
        if curve_points is not None:
            self.curve_points = int(curve_points)
        self.t = np.linspace(-np.pi, np.pi, self.curve_points)
        if self.n_bodies:
            self._upload()

    def load_conf(self, conf):
        self.gc = conf["gc"]
        self.coi_coef = conf["coi_coef"]

    # phys_body.py:177-198
    def load(self, conf, body_data, body_orb_data):
        self.load_conf(conf)
        self.mass = np.array(body_data["mass"], dtype=np.float64)
        self.a = np.array(body_orb_data["a"], dtype=np.float64)
        self.ecc = np.array(body_orb_data["ecc"], dtype=np.float64)
        self.pea = np.array(body_orb_data["pe_arg"], dtype=np.float64)
        self.ma = np.array(body_orb_data["ma"], dtype=np.float64)
        self.ref = np.array(body_orb_data["ref"], dtype=np.int64)
        self.dr = np.array(body_orb_data["dir"], dtype=np.float64)
        self.n_bodies = len(self.mass)
        self.pos = np.zeros((self.n_bodies, 2))
        self.ea = np.zeros(self.n_bodies)

    def _upload(self):
        self.dev.kepler_load(KIND_BODY, self.curve_points, self.a, self.b, self.f, self.ecc,
                             self.pea, self.n, self.dr, self.ma, self.ea, self.ref, self.t)
        self.dev.kepler_set(KIND_BODY, K_POS, self.pos)
        self.dev.kepler_set_order(np.argsort(self.coi)[-1::-1])    # phys_body.py:328

    # phys_body.py:221-312 (orbit part)
    def initial(self, warp):
        orb = calc_orb("body", self.ref, self.mass, self.mass, self.gc, self.coi_coef, self.a,
                       self.ecc)
        self.b, self.f, self.coi, self.pe_d, self.ap_d, self.period, self.n, self.u = (
            orb[k] for k in ("b", "f", "coi", "pe_d", "ap_d", "period", "n", "u"))
        body_orb = {"a": self.a, "b": self.b, "f": self.f, "coi": self.coi, "ref": self.ref,
                    "ecc": self.ecc, "pe_d": self.pe_d, "ap_d": self.ap_d, "pea": self.pea,
                    "n": self.n, "dir": self.dr, "per": self.period, "u": self.u}
        self._upload()
        self.move(warp)
        curves_mov = self.curve_move()
        return body_orb, self.pos, self.ma, self.ea, curves_mov

    # phys_body.py:323-346 -> (pos (N,2), ma (N,), ea (N,)), refreshed in place
    def move(self, warp):
        self.dev.kepler_move(KIND_BODY, warp)
        n = self.n_bodies
        self.dev.kepler_get(KIND_BODY, K_POS, n, out=self.pos)
        self.dev.kepler_get(KIND_BODY, K_MA, n, out=self.ma)
        self.dev.kepler_get(KIND_BODY, K_EA, n, out=self.ea)
        return self.pos, self.ma, self.ea

    # phys_body.py:315-320: relative curve of one body, (P,2)
    def curve(self, body):
        return self.dev.host_curve_points(self.ecc[body:body + 1], self.a[body:body + 1],
                                          self.b[body:body + 1], self.pea[body:body + 1],
                                          self.t)[0]

    # phys_body.py:405-411 -> (N,P,2); `visible` limits the copy to the curves the
    # caller draws (game.py:1403-1407), the full set stays resident on the device
    def curve_move(self, visible=None):
        self.dev.kepler_curves(KIND_BODY)
        if visible is None:
            return self.dev.kepler_curves_get(KIND_BODY, 0, self.n_bodies, self.curve_points)
        got = self.dev.kepler_curves_gather(KIND_BODY, visible, self.curve_points)
        return {int(i): got[k] for k, i in enumerate(visible)}

    # phys_body.py:201-218; radius / atm_h are the caller's body data (UI side)
    def culling(self, sim_screen, zoom, radius=None, atm_h=None, screen_x=1366):
        n = self.n_bodies
        radius = np.zeros(n) if radius is None else np.asarray(radius, dtype=np.float64)
        atm_h = np.zeros(n) if atm_h is None else np.asarray(atm_h, dtype=np.float64)
        dev = self.dev
        visible_bodies = dev.kepler_culling(KIND_BODY, n, (radius + atm_h) * zoom, sim_screen, zoom)
        coi_idx = dev.kepler_culling(KIND_BODY, n, self.coi, sim_screen, zoom)
        visible_coi = np.concatenate(([0], coi_idx))
        coi_truth = np.zeros(n, dtype=np.uint8)
        coi_truth[coi_idx] = 1
        self.visible_coi_truth = coi_truth
        visible_orbits = dev.kepler_visible_orbits(KIND_BODY, n, coi_truth, self.coi, self.a, zoom,
                                                   screen_x)
        return visible_bodies, visible_coi, visible_orbits
