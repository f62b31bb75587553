"""Drop-in for the hot path of volatilespace/physics/phys_body.py on a B200.

`Physics.load / initial / move / curve / curve_move / culling` keep the reference's names,
argument meaning and return shapes (phys_body.py:177-346, 405-411); the Kepler solve, position
compose, curve generation, culling and the body-data pass of `initial()` (thermal, colour,
classification -- phys_body.py:227-282) run in libvsb200.so.
"""
import numpy as np

from . import _lib
from ._lib import KIND_BODY, K_EA, K_MA, K_POS
from .device import Device, default_device, pinned_empty
from .phys_editor import STELLAR_CLASSES, body_radius, thermal_consts
from .synth import calc_orb


class Physics:
    """Body physics class (game mode)."""

    def __init__(self, device=None, curve_points=300):
        self.dev = device if isinstance(device, Device) else (
            default_device() if device is None else Device(device))
        self.curve_points = int(curve_points)
        self.t = np.linspace(-np.pi, np.pi, self.curve_points)   # phys_body.py:165
        self.gc, self.coi_coef = 1.0, 0.7
        self.rad_mult, self.mass_thermal_mult, self.min_mass = 179.0, 2e26, 200
        self.screen_x, self.screen_y = 1366, 768       # pygame surface size in the reference
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
        self.rad_mult = conf.get("rad_mult", self.rad_mult)
        self.mass_thermal_mult = conf.get("mass_thermal_mult", self.mass_thermal_mult)
        self.min_mass = conf.get("min_planet_mass", self.min_mass)

    # phys_body.py:177-198
    def load(self, conf, body_data, body_orb_data):
        self.load_conf(conf)
        self.mass = np.array(body_data["mass"], dtype=np.float64)
        n = len(self.mass)
        self.names = body_data.get("name", np.array(["body%d" % i for i in range(n)]))
        self.den = np.array(body_data.get("den", np.full(n, 1000.0)), dtype=np.float64)
        self.base_color = np.array(body_data.get("color", np.zeros((n, 3), int)),
                                   dtype=int).reshape(n, 3)
        self.atm_pres0 = np.array(body_data.get("atm_pres0", np.zeros(n)), dtype=np.float64)
        self.atm_scale_h = np.array(body_data.get("atm_scale_h", np.zeros(n)), dtype=np.float64)
        self.atm_den0 = np.array(body_data.get("atm_den0", np.zeros(n)), dtype=np.float64)
        self.a = np.array(body_orb_data["a"], dtype=np.float64)
        self.ecc = np.array(body_orb_data["ecc"], dtype=np.float64)
        self.pea = np.array(body_orb_data["pe_arg"], dtype=np.float64)
        self.ma = pinned_empty(len(self.mass))      # per-tick results in page-locked memory
        self.ma[:] = body_orb_data["ma"]
        self.ref = np.array(body_orb_data["ref"], dtype=np.int64)
        self.dr = np.array(body_orb_data["dir"], dtype=np.float64)
        self.n_bodies = len(self.mass)
        self.pos = pinned_empty((self.n_bodies, 2))
        self.pos[:] = 0.0
        self.ea = pinned_empty(self.n_bodies)
        self.ea[:] = 0.0

    def _upload(self):
        self.dev.kepler_load(KIND_BODY, self.curve_points, self.a, self.b, self.f, self.ecc,
                             self.pea, self.n, self.dr, self.ma, self.ea, self.ref, self.t)
        self.dev.kepler_set(KIND_BODY, K_POS, self.pos)
        self.dev.kepler_set_order(np.argsort(self.coi)[-1::-1])    # phys_body.py:328

    # phys_body.py:221-312 -> (body_data, body_orb, pos, ma, ea, curves_mov)
    def initial(self, warp):
        # body data, :227-282: radius on the host (np.cbrt), the rest in one device pass.  The
        # game-mode formulas equal the editor's with gc = 1 in surf_grav and atm_h starting at 0.
        n = self.n_bodies
        self.radius = body_radius(self.mass, self.den, self.rad_mult)
        th = self.dev.host_body_thermal(
            thermal_consts(1.0, self.rad_mult, self.mass_thermal_mult, self.min_mass), self.mass,
            self.den, self.radius, self.atm_pres0, self.atm_scale_h, self.atm_den0, np.zeros(n),
            self.base_color)
        self.atm_h, self.types = th["atm_h"], th["types"]
        body_data = {"name": self.names, "type": self.types, "mass": self.mass, "den": self.den,
                     "temp": th["temp"], "lum": th["luminosity"],
                     "class": STELLAR_CLASSES[th["stellar_class"]], "color_b": self.base_color,
                     "color": th["color"], "radius": self.radius, "rad_sc": th["rad_sc"],
                     "surf_grav": th["surf_grav"], "atm_pres0": self.atm_pres0,
                     "atm_scale_h": self.atm_scale_h, "atm_den0": self.atm_den0,
                     "atm_h": self.atm_h}
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
        return body_data, body_orb, self.pos, self.ma, self.ea, curves_mov

    # phys_body.py:323-346 -> (pos (N,2), ma (N,), ea (N,)), refreshed in place
    def move(self, warp):
        self.dev.kepler_move_state(KIND_BODY, warp, None, self.pos, self.ma, self.ea)
        return self.pos, self.ma, self.ea

    # phys_body.py:315-320: relative curve of one body, (P,2)
    def curve(self, body):
        return self.dev.host_curve_points(self.ecc[body:body + 1], self.a[body:body + 1],
                                          self.b[body:body + 1], self.pea[body:body + 1],
                                          self.t)[0]

    # phys_body.py:405-411 -> (N,P,2); `visible` limits the copy to the curves the
    # caller draws (game.py:1403-1407), the full set stays resident on the device
    def curve_move(self, visible=None):
        if visible is None:
            return self.dev.kepler_curve_move(KIND_BODY, self.n_bodies, self.curve_points)
        self.dev.kepler_curves(KIND_BODY)
        got = self.dev.kepler_curves_gather(KIND_BODY, visible, self.curve_points, reuse=True)
        return {int(i): got[k] for k, i in enumerate(visible)}

    # phys_body.py:201-218
    def culling(self, sim_screen, zoom):
        n = self.n_bodies
        dev = self.dev
        visible_bodies = dev.kepler_culling(KIND_BODY, n, (self.radius + self.atm_h) * zoom,
                                            sim_screen, zoom)
        coi_idx = dev.kepler_culling(KIND_BODY, n, self.coi, sim_screen, zoom)
        visible_coi = np.concatenate(([0], coi_idx))
        coi_truth = np.zeros(n, dtype=np.uint8)
        coi_truth[coi_idx] = 1
        visible_orbits = dev.kepler_visible_orbits(KIND_BODY, n, coi_truth, self.coi, self.a, zoom,
                                                   self.screen_x)
        return visible_bodies, visible_coi, visible_orbits
