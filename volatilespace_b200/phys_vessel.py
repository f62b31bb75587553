"""Drop-in for the physics of volatilespace/physics/phys_vessel.py on a B200:
`Physics.move / curve / curve_move` (phys_vessel.py:477-503, 364-369) and the orbit
points `Physics.points` (:372-474: body impact, atmosphere entry, COI leave, COI enter
-- SURVEY 8f N2), batched over all vessels, and the per-tick event scans built on those
points: `check_warp` (:506-560), `predict_enter_coi_service` (:744-758), `cross_coi`
(:563-632) and `change_vessel` (:306-361), and the drawing-side `curve_segments` (:828-1075:
the lit / dark part of every visible orbit curve, cut on the device from the resident curves);
the one-vessel-per-frame helpers `selected`, `rotate`, `predict_next_orbit` are host code
(vessel_draw.py).
"""
import numpy as np

from ._lib import KIND_VESSEL, K_EA, K_MA, K_POS
from .device import Device, default_device, pinned_empty
from . import vessel_draw
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
        # after initial() the orbit points, the hold mask and prev_ea live on the device next to
        # the vessel elements, and the per-tick scans read them there (nothing per-vessel over
        # PCIe); before that (or with resident_events = False) every call ships its arrays
        self.resident_events = True
        self._resident = False

    # phys_vessel.py:198-240 (orbit arrays only)
    def load(self, conf, body_data, body_orb, vessel_data, vessel_orb_data):
        self.gc = conf["gc"]
        self.body_mass = np.array(body_data["mass"], dtype=np.float64)
        body_orb = body_orb or {}
        nb = len(self.body_mass)
        self.body_coi = np.array(body_orb.get("coi", np.zeros(nb)), dtype=np.float64)
        # what points() needs of the bodies (phys_vessel.py:200-213); absent keys -> zeros
        col = lambda d, k: np.array((d or {}).get(k, np.zeros(nb)), dtype=np.float64)
        self.body_size = col(body_data, "radius")
        self.body_atm = col(body_data, "atm_h")
        self.body_ref = np.array(body_orb.get("ref", np.zeros(nb)), dtype=np.int64)
        self.body_a, self.body_b, self.body_f, self.body_ecc, self.body_pea, self.body_n, \
            self.body_dr = (col(body_orb, k) for k in ("a", "b", "f", "ecc", "pea", "n", "dir"))
        self.predict_coi_limit = 300                   # [game] predict_coi_limit, :181
        self.entered_coi = self.left_coi = self.left_coi_prev_ref = None
        self._hold = np.zeros(0, dtype=np.uint8)       # membership mask of physical_hold
        self.names = np.array((vessel_data or {}).get("name", []))
        self.a = np.array(vessel_orb_data["a"], dtype=np.float64)
        self.ecc = np.array(vessel_orb_data["ecc"], dtype=np.float64)
        self.pea = np.array(vessel_orb_data["pe_arg"], dtype=np.float64)
        self.ref = np.array(vessel_orb_data["ref"], dtype=np.int64)
        self.dr = np.array(vessel_orb_data["dir"], dtype=np.float64)
        self.n_vessels = len(self.a)
        # the per-tick results (pos, ma, ea, refreshed in place by move()) live in page-locked
        # memory: their device -> host copies then run at PCIe rate
        self.ma = pinned_empty(self.n_vessels)
        self.ma[:] = vessel_orb_data["ma"]
        zeros = np.zeros(self.n_vessels)
        self.rot_angle = np.array((vessel_data or {}).get("rot_angle", zeros), dtype=np.float64)  # :216-218
        self.rot_acc = np.array((vessel_data or {}).get("rot_acc", zeros), dtype=np.float64)
        self.rot_speed = np.zeros(self.n_vessels)
        self.intersect_time = np.zeros(self.n_vessels)
        self.pos = pinned_empty((self.n_vessels, 2))
        self.pos[:] = 0.0
        self.ea = pinned_empty(self.n_vessels)
        self.ea[:] = vessel_orb_data.get("ea", 0.0)
        self._hold = np.zeros(self.n_vessels, dtype=np.uint8)
        self.prev_ea = np.array(self.ea)
        nan = lambda w: np.full((self.n_vessels, w), np.nan)
        self.body_impact, self.body_enter_atm, self.coi_leave, self.coi_enter = (
            nan(2), nan(2), nan(2), nan(6))            # :231-234

    # phys_vessel.py:276-303 (orbit part) -> (vessel_orb, pos, ma, curves_mov)
    def initial(self, warp, body_pos, body_ma=None, body_ea=None):
        orb = calc_orb("vessel", self.ref, self.body_mass, None, self.gc, 0.0, self.a, self.ecc)
        self.b, self.f, self.pe_d, self.ap_d, self.period, self.n, self.u = (
            orb[k] for k in ("b", "f", "pe_d", "ap_d", "period", "n", "u"))
        vessel_orb = {"a": self.a, "ref": np.array(self.ref), "ecc": self.ecc, "pe_d": self.pe_d,
                      "ap_d": self.ap_d, "pea": self.pea, "dir": self.dr, "per": self.period}
        self.dev.kepler_load(KIND_VESSEL, self.curve_points, self.a, self.b, self.f, self.ecc,
                             self.pea, self.n, self.dr, self.ma, self.ea, self.ref, self.t)
        self._resident = bool(self.resident_events and self.n_vessels)
        if self._resident:
            self.dev.vessel_events_load(self.pe_d, self.ap_d, self.period, self.body_impact,
                                        self.body_enter_atm, self.coi_leave, self.coi_enter,
                                        self._hold, self.prev_ea)
        self.move(warp, body_pos, body_ma, body_ea)
        curves_mov = self.curve_move()
        if body_ma is not None and body_ea is not None and self.n_vessels:
            self.points()                              # :300-302, all vessels in one batch
        return vessel_orb, self.pos, self.ma, curves_mov

    # phys_vessel.py:477-494 -> (pos (Nv,2), ma (Nv,)), refreshed in place
    def move(self, warp, body_pos, body_ma=None, body_ea=None):
        self.body_pos = body_pos
        if body_ma is not None:
            self.body_ma = np.asarray(body_ma, dtype=np.float64)
        if body_ea is not None:
            self.body_ea = np.asarray(body_ea, dtype=np.float64)
        self.prev_ea = np.array(self.ea)
        self.dev.kepler_move_state(KIND_VESSEL, warp, body_pos, self.pos, self.ma, self.ea)
        return self.pos, self.ma

    # phys_vessel.py:372-474; vessel=None processes every vessel (the loop of initial())
    def points(self, vessel=None, only_enter_coi=False):
        """Characteristic points of the vessel orbits: body_impact, body_enter_atm, coi_leave
        (each (Nv,2): next / previous eccentric anomaly) and coi_enter (Nv,6), updated in place."""
        self._points(None if vessel is None else [vessel], only_enter_coi)

    def _points(self, sel, only_enter_coi=False):
        body12 = np.stack([self.body_a, self.body_b, self.body_f, self.body_ecc, self.body_ma,
                           self.body_ea, self.body_pea, self.body_n, self.body_dr, self.body_coi,
                           self.body_size, self.body_atm], axis=1)
        kw = dict(sel=sel, entered_coi=self.entered_coi, left_coi=self.left_coi,
                  left_coi_prev_ref=self.left_coi_prev_ref, only_enter_coi=only_enter_coi,
                  steps_limit=self.predict_coi_limit)
        if self._resident:
            self.dev.vessel_points(body12, self.body_ref, **kw)
            got = self.dev.vessel_events_get(self.n_vessels)
            for key in ("body_impact", "body_enter_atm", "coi_leave", "coi_enter"):
                getattr(self, key)[:] = got[key]       # host mirrors, refreshed in place
            return
        vessel12 = np.stack([self.a, self.b, self.f, self.ecc, self.ma, self.ea, self.pea, self.n,
                             self.dr, self.pe_d, self.ap_d, self.period], axis=1)
        self.dev.host_vessel_points(vessel12, self.ref, body12, self.body_ref, self.body_impact,
                                    self.body_enter_atm, self.coi_leave, self.coi_enter, **kw)

    @property
    def physical_hold(self):
        """Vessels held in physical warp (the reference keeps them in append order; ascending
        here -- only membership and the count are ever used, game.py:1171-1215)."""
        if self._resident:
            self._hold[:] = self.dev.vessel_events_get(self.n_vessels)["hold"]
        return np.flatnonzero(self._hold)

    # phys_vessel.py:506-560 -> (situation, alarm_vessel, len(physical_hold))
    def check_warp(self, warp):
        if self._resident:
            return self.dev.vessel_check_warp(warp)
        return self.dev.host_check_warp(self.ecc, self.dr, self.ea, self.ma, self.n,
                                        self.body_impact, self.body_enter_atm, self.coi_leave,
                                        self.coi_enter, warp, self._hold)

    # phys_vessel.py:744-758
    def predict_enter_coi_service(self):
        pending = self.left_coi is not None or self.entered_coi is not None
        if self._resident:
            again = self.dev.vessel_enter_coi_service(self.n_vessels, pending)
        else:
            again = self.dev.host_enter_coi_service(self.dr, self.prev_ea, self.ea, self.coi_enter,
                                                    pending)
        if len(again):
            self._points(again, only_enter_coi=True)
        return again

    # phys_vessel.py:563-632 -> the vessel whose orbit changed, or None
    def cross_coi(self):
        if self._resident:
            ves, kind = self.dev.vessel_cross_scan()
        else:
            ves, kind = self.dev.host_cross_scan(self.ecc, self.dr, self.prev_ea, self.ea,
                                                 self.body_impact, self.coi_leave, self.coi_enter)
        # the reference's loop clears these flags for every vessel it visits (:579-583): up to and
        # including the crossing one, or all of them
        last = self.n_vessels - 1 if ves is None else ves
        if self.entered_coi is not None and self.entered_coi <= last:
            self.entered_coi = None
        if self.left_coi is not None and self.left_coi <= last:
            self.left_coi = None
            self.left_coi_prev_ref = None
        if ves is None:
            return None
        ref = int(self.ref[ves])
        new_ref = int(self.body_ref[ref]) if kind == 2 else int(self.coi_enter[ves, 0])
        ea_cross = self.coi_leave[ves, 0] if kind == 2 else self.coi_enter[ves, 1]
        body7 = np.stack([self.body_a, self.body_ecc, self.body_pea, self.body_dr, self.body_mass,
                          self.body_pos[:, 0], self.body_pos[:, 1]], axis=1)
        row = [self.a[ves], self.b[ves], self.f[ves], self.ecc[ves], self.pea[ves], self.u[ves],
               self.dr[ves], ea_cross]
        out, status = self.dev.host_cross_apply(kind, row, ref, new_ref, body7, self.gc)
        if status[0]:
            return None                                # :609-610
        self.a[ves], self.ecc[ves], self.pea[ves], self.ma[ves], self.dr[ves] = out[0, :5]
        self.ref[ves] = new_ref
        if kind == 2:
            self.left_coi, self.left_coi_prev_ref = ves, ref
        else:
            self.entered_coi = ves
        return ves

    # phys_vessel.py:306-361 -> (vessel_data, vessel_orb)
    def change_vessel(self, vessel):
        sl = slice(vessel, vessel + 1)
        orb = calc_orb("vessel", self.ref[sl], self.body_mass, None, self.gc, 0.0, self.a[sl],
                       self.ecc[sl])
        for key, arr in (("b", self.b), ("f", self.f), ("pe_d", self.pe_d), ("ap_d", self.ap_d),
                         ("period", self.period), ("n", self.n), ("u", self.u)):
            arr[vessel] = orb[key][0]
        ecc, ma = self.ecc[sl], self.ma[sl]
        if ecc[0] < 1:
            self.ea[vessel] = self.dev.host_solve_kepler_ell(ecc, ma, 1e-10)[0]
        else:
            self.ea[vessel] = self.dev.host_newton_root_kepler_hyp(ecc, ma, ma)[0]
        self._push_vessel(vessel)          # curve(vessel): the device regenerates it from the elements
        self.points(vessel)
        vessel_orb = [self.a[vessel], self.ref[vessel], self.ecc[vessel], self.pe_d[vessel],
                      self.ap_d[vessel], self.pea[vessel], self.dr[vessel], self.period[vessel]]
        return [self.names[vessel] if len(self.names) > vessel else None], vessel_orb

    def _push_vessel(self, vessel):
        """Write one vessel's (changed) elements into the device-resident vessel state."""
        from . import _lib as L
        for field, arr in ((L.K_A, self.a), (L.K_B, self.b), (L.K_F, self.f), (L.K_ECC, self.ecc),
                           (L.K_PEA, self.pea), (L.K_N, self.n), (L.K_DR, self.dr), (L.K_MA, self.ma),
                           (L.K_EA, self.ea), (L.K_REF, self.ref)):
            self.dev.kepler_set(KIND_VESSEL, field, arr[vessel:vessel + 1], first=vessel)
        if self._resident:
            self.dev.vessel_events_set_orbit(vessel, self.pe_d[vessel], self.ap_d[vessel],
                                             self.period[vessel])

    # phys_vessel.py:364-369
    def curve(self, vessel):
        return self.dev.host_curve_points(self.ecc[vessel:vessel + 1], self.a[vessel:vessel + 1],
                                          self.b[vessel:vessel + 1], self.pea[vessel:vessel + 1],
                                          self.t)[0]

    # phys_vessel.py:497-503
    def curve_move(self, visible=None):
        if visible is None:
            return self.dev.kepler_curve_move(KIND_VESSEL, self.n_vessels, self.curve_points)
        self.dev.kepler_curves(KIND_VESSEL)
        got = self.dev.kepler_curves_gather(KIND_VESSEL, visible, self.curve_points, reuse=True)
        return {int(i): got[k] for k, i in enumerate(visible)}

    # phys_vessel.py:258-273
    def culling(self, sim_screen, zoom, visible_coi):
        n = self.n_vessels
        dev = self.dev
        self.visible_vessels = dev.kepler_culling(KIND_VESSEL, n, 10 / zoom, sim_screen, zoom)
        coi_truth = np.zeros(len(self.body_coi), dtype=np.uint8)
        coi_truth[np.asarray(visible_coi, dtype=np.int64)] = 1
        self.visible_orbits = dev.kepler_visible_orbits(KIND_VESSEL, n, coi_truth, self.body_coi,
                                                        self.pe_d, zoom, self.screen_x)
        return self.visible_vessels, self.visible_orbits

    # phys_vessel.py:828-1075
    def curve_segments(self, dense=None):
        """Two segments for each curve on screen (after curve_move() and points()): returns
        (curves_light, curves_dark, first_intersect, intersect_type, intersect_time,
        select_range) like the reference.  Rows of vessels outside `visible_orbits` -- np.empty
        garbage in the reference -- are NaN; `curves_dark` is also NaN for a visible vessel
        without any orbit point (the reference leaves that row unwritten).  With dense=False (the
        default above 65 536 vessels) the two curve sets are dicts {vessel: (P,2) array} instead of
        (Nv,P,2) arrays, so that only visible curves ever exist on the host."""
        n, P = self.n_vessels, self.curve_points
        vis = np.asarray(self.visible_orbits, dtype=np.int64)
        if not self._resident:
            self.dev.vessel_events_load(self.pe_d, self.ap_d, self.period, self.body_impact,
                                        self.body_enter_atm, self.coi_leave, self.coi_enter,
                                        self._hold, self.prev_ea)
        light, dark, first, itype, itime, srange = self.dev.vessel_curve_segments(vis, P, reuse=True)
        first_intersect = np.full((n, 2), np.nan)
        select_range = np.full((n, 2), np.nan)
        intersect_type = np.zeros(n, dtype=np.int64)
        intersect_time = np.zeros(n)
        first_intersect[vis] = first
        select_range[vis] = srange
        intersect_type[vis] = itype
        intersect_time[vis] = itime
        self.intersect_time = intersect_time
        if dense is None:
            dense = n <= 65536
        if dense:
            curves_light = np.full((n, P, 2), np.nan)
            curves_dark = np.full((n, P, 2), np.nan)
            curves_light[vis] = light
            curves_dark[vis] = dark
        else:
            curves_light = {int(v): light[k] for k, v in enumerate(vis)}
            curves_dark = {int(v): dark[k] for k, v in enumerate(vis)}
        return (curves_light, curves_dark, first_intersect, intersect_type, intersect_time,
                select_range)

    # phys_vessel.py:777-825
    def selected(self, vessel):
        """Physics for the selected vessel (after move()): (ta, pe, pe_t, ap, ap_t, distance,
        speed_orb, speed_hor, speed_vert).  One vessel per frame: host scalar math."""
        el = {k: float(getattr(self, k)[vessel]) for k in
              ("a", "b", "ecc", "pea", "n", "dr", "ma", "ea", "u", "pe_d", "ap_d")}
        return vessel_draw.selected(el, self.pos[vessel], np.asarray(self.body_pos)[self.ref[vessel]])

    # phys_vessel.py:761-774
    def rotate(self, warp, vessel, direction):
        self.rot_angle = vessel_draw.rotate(warp, vessel, direction, self.rot_angle, self.rot_speed,
                                            self.rot_acc)
        return self.rot_angle

    # phys_vessel.py:635-741
    def predict_next_orbit(self, vessel):
        """Next orbit of `vessel` after its coming COI change (orbit line and ghost body): the
        reference's 11-tuple, all None when there is nothing ahead.  Needs curve_segments()."""
        return vessel_draw.predict_next_orbit(self, vessel)

    def tick(self, warp, body_pos):
        """move + curve_move in one fused kernel (results stay on the device)."""
        self.dev.kepler_tick(KIND_VESSEL, warp, body_pos)
