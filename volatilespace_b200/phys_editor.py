"""Drop-in for volatilespace/physics/phys_editor.py's hot path on a B200.

`Physics` keeps the reference's method names, argument meaning and return
conventions (phys_editor.py:112-857) so editor.py can bind
`physics = volatilespace_b200.phys_editor.Physics()` instead of the numba class;
the O(N^2) loops (find_parents, check_collision, simplified_orbit_coi) and the
per-tick O(N) kernels run in libvsb200.so.  `AllPairsPhysics` is the north-star
MODE_ALLPAIRS variant (documentation/orbital-physics/01-complex-orbit-model.txt).

Host numpy arrays with the reference's attribute names mirror the device state:
hot calls refresh them IN PLACE (callers hold aliases between calls,
editor.py:1517) -- eagerly by default, or on `pull()` with sync="lazy".
There is no CPU fallback: constructing a Physics without the CUDA library or
without a GPU raises.
"""
import math

import numpy as np

from . import _lib
from ._lib import (F_ACC, F_COI, F_DEN, F_ECC_V, F_FOCUS, F_MASS, F_PARENTS, F_PERIAPSIS_ARG,
                   F_POS, F_RAD, F_REL_VEL, F_SEMI_MAJOR, F_SEMI_MINOR, F_VEL)
from .device import Device, default_device, pinned_empty, scratch_device

_FIELDS = {"mass": F_MASS, "den": F_DEN, "rad": F_RAD, "coi": F_COI, "pos": F_POS, "vel": F_VEL,
           "rel_vel": F_REL_VEL, "parents": F_PARENTS, "focus": F_FOCUS, "ecc_v": F_ECC_V,
           "semi_major": F_SEMI_MAJOR, "semi_minor": F_SEMI_MINOR,
           "periapsis_arg": F_PERIAPSIS_ARG}
_WIDE = ("pos", "vel", "rel_vel", "ecc_v")

DEFAULT_CONF = {"gc": 1.0, "rad_mult": 179.0, "mass_thermal_mult": 2e26, "coi_coef": 0.7,
                "vessel_scale": 0.1, "min_planet_mass": 200}   # defaults.py:58-65


# phys_shared.py:13-17, same expressions so the constants carry the same roundings
GC_REAL = 6.674 * 10**-11
C_LIGHT = 299792458
SIGMA = 5.670374419 * 10**-8
LS = 3.828 * 10**26
MS = 1.9884 * 10**30
STELLAR_CLASSES = np.array(["Unknown", "M", "K", "G", "F", "A", "B", "O"])


def thermal_consts(gc_sim, rad_mult, mass_thermal_mult, min_mass):
    return np.array([gc_sim, rad_mult, mass_thermal_mult, min_mass, GC_REAL, C_LIGHT, LS, MS,
                     np.sqrt(2), np.pi**(1 / 4), SIGMA**(1 / 4), 4 * np.pi], dtype=np.float64)


def _ellipse_velocity_angle(pr, a, b, f, omega):
    """Direction of motion at the focus-relative point `pr` of an ellipse (semi-axes a, b, focus
    distance f) rotated by omega: slope from the implicit derivative of the rotated conic, and the
    half-plane test against the chord through the curve's x-extremes that tells the two branches
    apart (phys_editor.py:486-503; the construction of convert.kepler_to_velocity :24-43)."""
    so, co = math.sin(omega), math.cos(omega)
    a2, b2 = a**2, b**2
    cx, cy = pr[0] - f * np.cos(omega), pr[1] - f * np.sin(omega)      # relative to the centre
    angle = np.arctan(-(b2 * cx * co**2 + a2 * cx * so**2 + b2 * cy * so * co - a2 * cy * so * co)
                      / (a2 * cy * co**2 + b2 * cy * so**2 + b2 * cx * so * co - a2 * cx * so * co))
    x_max = math.sqrt(a2 * math.cos(2 * omega) + a2 - b2 * math.cos(2 * omega) + b2) / math.sqrt(2) - 10**-6
    x2, c2, s2 = x_max**2, co**2, so**2
    c4, s4 = (co * co) * (co * co), (so * so) * (so * so)
    y_max = (a * b * math.sqrt(-x2 * (c4 + s4) - 2 * x2 * c2 * s2 + b2 * s2 + a2 * c2)
             + a2 * x_max * so * co - b2 * x_max * so * co) / (a2 * c2 + b2 * s2)
    x1, y1 = x_max + f * co, y_max + f * so
    x2_, y2_ = -x_max + f * co, -y_max + f * so
    if ((x2_ - x1) * (pr[1] - y1)) - ((y2_ - y1) * (pr[0] - x1)) < 0:
        angle += np.pi
    return angle % (2 * np.pi)


def mass_order(mass):
    """np.argsort(mass)[-1::-1], phys_editor.py:353,375 (host numpy)."""
    return np.ascontiguousarray(np.argsort(mass)[-1::-1], dtype=np.int64)


def body_radius(mass, den, rad_mult):
    """phys_editor.py:580-581; stays on the host so `rad` is bit-identical to the
    reference's np.cbrt path (it feeds the bit-exact collision predicate)."""
    volume = mass / den
    return rad_mult * np.cbrt(3 * volume / (4 * np.pi))


# ---- the per-body free functions of the reference module (phys_editor.py:36-97), for callers
# that import them directly.  Each evaluates the whole batch on the device; the `_one` forms pick
# one body out of it and are meant for tests and occasional queries, not for loops.
def find_parents_all(bodies_sorted, pos, coi, device=None):
    """find_parent_one for every body -> parents (N,) int64."""
    dev = device or scratch_device()
    return dev.host_find_parents(bodies_sorted, pos, coi)


def find_parent_one(body, bodies_sorted, pos, coi, device=None):
    """phys_editor.py:36-48."""
    return int(find_parents_all(bodies_sorted, pos, coi, device)[body])


def gravity_one(body, parent, mass, rel_pos, gc, device=None):
    """phys_editor.py:51-60: acceleration of `body` toward `parent` (rel_pos = pos[parent] -
    pos[body]); the root (body 0) gets zeros.  Evaluated by the all-pairs kernel on the two-body
    system {body at the origin, parent at rel_pos} -- tolerance-level agreement (rsqrt^3 instead
    of sqrt / atan2 / cos / sin)."""
    if not body:
        return np.zeros(2)
    dev = device or scratch_device()
    pos = np.array([[0.0, 0.0], [rel_pos[0], rel_pos[1]]])
    return dev.host_allpairs_accel(np.array([mass[body], mass[parent]], dtype=np.float64), pos, gc)[0]


def kepler_basic_all(parents, mass, pos, vel, gc, coi_coef, device=None):
    """kepler_basic_one for every body -> (focus, ecc_v (N,2), semi_major, semi_minor,
    periapsis_arg, coi)."""
    dev = device or scratch_device()
    par = _lib.i64(parents)
    mass, pos, vel = _lib.f64(mass), _lib.f64(pos), _lib.f64(vel)
    n = len(mass)
    focus, sma, smi, pea, coi = (np.empty(n) for _ in range(5))
    ecc_v = np.empty((n, 2))
    _lib.check(dev.L.vsb_host_kepler_basic(dev.h, n, _lib.ptr(par), _lib.ptr(mass), _lib.ptr(pos),
                                           _lib.ptr(vel), float(gc), float(coi_coef),
                                           _lib.ptr(focus), _lib.ptr(ecc_v), _lib.ptr(sma),
                                           _lib.ptr(smi), _lib.ptr(pea), _lib.ptr(coi)))
    return focus, ecc_v, sma, smi, pea, coi


def kepler_basic_one(body, parent, mass, pos, vel, gc, coi_coef, device=None):
    """phys_editor.py:63-83 -> (focus, ecc_v, semi_major, semi_minor, periapsis_arg, coi)."""
    if not body:
        return 0.0, np.zeros(2), 0.0, 0.0, 0.0, 0.0
    # two-body system: row 0 = the parent (acts as root), row 1 = the body
    out = kepler_basic_all([0, 0], [mass[parent], mass[body]], [pos[parent], pos[body]],
                           [vel[parent], vel[body]], gc, coi_coef, device)
    return tuple(o[1] for o in out)


def check_collision_one(body1, mass, pos, rad, device=None):
    """phys_editor.py:86-97: first body2 > body1 touching body1 -> (body_del, body_add) or None."""
    dev = device or scratch_device()
    mass, pos, rad = (np.asarray(x, dtype=np.float64) for x in (mass, pos, rad))
    if body1 >= len(mass) - 1:
        return None
    # rows body1.. : a pair (0, j) is the lexicographic minimum whenever body1 collides at all
    d, a = dev.host_check_collision(mass[body1:], pos[body1:], rad[body1:])
    if d is None or min(d, a) != 0:
        return None
    return d + body1, a + body1


class _DeviceBodies:
    """Shared plumbing: host mirrors <-> device fields."""

    def __init__(self, device=None, sync="eager"):
        assert sync in ("eager", "lazy")
        self.dev = device if isinstance(device, Device) else (
            default_device() if device is None else Device(device))
        self.sync = sync
        self._stale = set()        # host mirrors older than the device copy

    def _alloc_host(self, n):
        """Host mirrors; from 8192 rows up they are page-locked (vsb_host_register), so that
        pull() / _push() move them at PCIe rate."""
        for name in _FIELDS:
            shape = (n, 2) if name in _WIDE else (n,)
            arr = pinned_empty(shape, np.int64 if name == "parents" else np.float64, self.dev.L)
            arr[...] = 0
            setattr(self, name, arr)

    def _push(self, *names):
        for name in names or _FIELDS:
            self.dev.set(_FIELDS[name], getattr(self, name))
            self._stale.discard(name)

    def pull(self, *names):
        """Refresh host mirrors in place from the device (one copy for several fields)."""
        names = names or tuple(self._stale)
        # small systems: one packed copy for all fields; large ones: straight into the page-locked
        # mirrors, field by field (no second pass over the data on the host)
        if len(names) > 1 and len(self.mass) <= 65536:
            n = len(self.mass)
            words = sum(n * (2 if k in _WIDE else 1) for k in names)
            buf = self.dev.get_many([_FIELDS[k] for k in names],
                                    out=self.dev.staging("pull", words))
            off = 0
            for k in names:
                w = 2 if k in _WIDE else 1
                chunk = buf[off:off + n * w]
                off += n * w
                dst = getattr(self, k)
                if k == "parents":
                    dst[:] = chunk.view(np.int64)
                else:
                    dst[...] = chunk.reshape(dst.shape)
                self._stale.discard(k)
            return
        for name in names:
            self.dev.get(_FIELDS[name], out=getattr(self, name))
            self._stale.discard(name)

    def _touched(self, *names):
        self._stale.update(names)
        if self.sync == "eager":
            self.pull(*names)

    def _row(self, name, i):
        """Current value of one row (device is authoritative when the mirror is stale)."""
        if name in self._stale:
            return self.dev.get(_FIELDS[name], first=int(i), count=1)[0]
        return getattr(self, name)[i]

    def _set_row(self, name, i, value):
        arr = getattr(self, name)
        if name not in self._stale:
            arr[i] = value
        v = np.ascontiguousarray(value, dtype=arr.dtype).reshape(-1)
        self.dev.set(_FIELDS[name], v, first=int(i))

    def _delete_host_row(self, i):
        for name in _FIELDS:
            setattr(self, name, np.delete(getattr(self, name), i, axis=0))


# host-only per-body data of phys_editor.Physics (:114-131) that the editor UI reads; they follow
# the structural operations exactly like there (set_root :293-313 swaps only _AUX_SWAPPED)
_AUX = ("temp", "luminosity", "stellar_class", "color", "base_color", "atm_h", "types", "rad_sc",
        "surf_grav", "atm_pres0", "atm_scale_h", "atm_den0")
_AUX_SWAPPED = ("temp", "luminosity", "stellar_class", "color", "base_color", "atm_h", "types",
                "rad_sc")
_AUX_DELETED = ("temp", "luminosity", "stellar_class", "color", "base_color", "atm_h", "types",
                "rad_sc", "atm_pres0", "atm_scale_h", "atm_den0")


class Physics(_DeviceBodies):
    """Editor physics class (MODE_REF: the reference's parent/COI model)."""

    def __init__(self, device=None, sync="eager", curve_points=300):
        super().__init__(device, sync)
        self.gc = DEFAULT_CONF["gc"]
        self.rad_mult = DEFAULT_CONF["rad_mult"]
        self.mass_thermal_mult = DEFAULT_CONF["mass_thermal_mult"]
        self.coi_coef = DEFAULT_CONF["coi_coef"]
        self.min_mass = DEFAULT_CONF["min_planet_mass"]
        self.largest = 0
        self.names = np.array([])
        self._order_cache = None
        self._order_on_device = False
        self._rad_dirty = True
        self._alloc_host(0)
        self._alloc_aux(0)
        self.reload_settings(curve_points)

    def _alloc_aux(self, n, body_data=None):
        """UI-side arrays, phys_editor.py:169-201."""
        bd = body_data or {}
        self.temp = np.zeros(n)
        self.luminosity = np.zeros(n)
        self.stellar_class = np.array(["Unknown"] * n)
        self.color = np.zeros((n, 3), int)
        self.base_color = np.array(bd.get("color", np.zeros((n, 3), int)), dtype=int).reshape(n, 3)
        self.atm_pres0 = np.array(bd.get("atm_pres0", np.zeros(n)), dtype=np.float64)
        self.atm_scale_h = np.array(bd.get("atm_scale_h", np.zeros(n)), dtype=np.float64)
        self.atm_den0 = np.array(bd.get("atm_den0", np.zeros(n)), dtype=np.float64)
        self.atm_h = np.zeros(n)
        self.types = np.zeros(n)
        self.rad_sc = np.zeros(n)
        self.surf_grav = np.zeros(n)

    # phys_editor.py:152-157
    def reload_settings(self, curve_points=None):
        if curve_points is not None:
            self.curve_points = int(curve_points)
        self.ell_t = np.linspace(-np.pi, np.pi, self.curve_points)
        self.par_t = np.linspace(-np.pi - 1, np.pi + 1, self.curve_points)

    # phys_editor.py:160-166
    def load_conf(self, conf):
        self.gc = conf["gc"]
        self.rad_mult = conf["rad_mult"]
        self.mass_thermal_mult = conf.get("mass_thermal_mult", self.mass_thermal_mult)
        self.coi_coef = conf["coi_coef"]
        self.min_mass = conf.get("min_planet_mass", self.min_mass)

    # phys_editor.py:169-201
    def load_system(self, conf, body_data, body_orb_data):
        self.load_conf(conf)
        mass = np.array(body_data["mass"], dtype=np.float64)
        n = len(mass)
        self.names = body_data.get("name", np.array(["body%d" % i for i in range(n)]))
        self._alloc_host(n)
        self._alloc_aux(n, body_data)
        self.mass[:] = mass
        self.den[:] = body_data["den"]
        self.rad[:] = body_radius(self.mass, self.den, self.rad_mult)
        self.pos[:] = body_orb_data["pos"]
        self.vel[:] = body_orb_data["vel"]
        self._stale.clear()
        self._mass_changed()
        self.dev.bodies_resize(n)
        self._push()
        self.simplified_orbit_coi()
        self.find_parents()
        self.pull("parents")
        self.rel_vel[:] = self.vel - self.vel[self.parents]
        self._push("rel_vel")
        self.body()

    # phys_editor.py:325-346
    def simplified_orbit_coi(self):
        self.largest = self.dev.simplified_orbit_coi(self.gc, self.coi_coef, self._order())
        self._touched("coi")

    # phys_editor.py:350-354
    def find_parents(self):
        self.dev.find_parents(self._order(), want=False)
        self._touched("parents")

    # phys_editor.py:357-378
    def gravity(self):
        order = None if self._order_on_device else self._order()
        if self._small_eager():
            # one launch: the fused kernel runs the tick and leaves the frame record in mapped
            # host memory (its collision answer is not used here)
            rec = np.empty(4 + 14 * len(self.mass))
            self.dev.editor_ticks(self.gc, self.coi_coef, order, 1, do_kepler_basic=False, frame_out=rec)
            self._order_on_device = True
            self._unpack_frame(rec, ("pos", "vel", "rel_vel", "parents"))
            return
        self.dev.gravity_ref(self.gc, order)
        self._order_on_device = True
        self._touched("parents", "rel_vel", "vel", "pos")

    def _small_eager(self):
        """Eager host mirrors of a system the fused one-CTA kernel serves (up to 1024 bodies): the
        per-method calls then take one launch each and read their results from mapped host memory
        instead of a kernel + a packing kernel + a copy."""
        return self.sync == "eager" and 0 < len(self.mass) <= 1024

    # phys_editor.py:577-599.  Everything here depends on mass, den and the atmosphere parameters
    # only, so it is recomputed when one of those changed (the reference recomputes it every tick)
    def body(self):
        if not self._rad_dirty:
            return
        rad = body_radius(self.mass, self.den, self.rad_mult)
        if not np.array_equal(rad, self.rad):
            self.rad[:] = rad
            self._push("rad")
        if len(self.mass):
            out = self.body_data(self.atm_pres0, self.atm_scale_h, self.atm_den0, self.atm_h,
                                 self.base_color)
            self.surf_grav, self.atm_h = out["surf_grav"], out["atm_h"]
            self.luminosity, self.temp, self.rad_sc = out["luminosity"], out["temp"], out["rad_sc"]
            self._color, self._types, self._stellar = (out["color"], out["types"],
                                                       out["stellar_class"])
        self._rad_dirty = False

    # phys_editor.py:602-651
    def temp_color(self):
        self.body()
        self.color = self._color
        return self.color

    # phys_editor.py:654-670
    def classify(self):
        self.body()
        self.types = self._types
        self.stellar_class = self._stellar

    # phys_editor.py:583-599 + temp_color :602-651 + classify :654-670: the
    # per-frame body data the editor UI shows, one device pass over host arrays
    def body_data(self, atm_pres0=None, atm_scale_h=None, atm_den0=None, atm_h=None,
                  base_color=None):
        n = len(self.mass)
        z = np.zeros(n)
        out = self.dev.host_body_thermal(
            thermal_consts(self.gc, self.rad_mult, self.mass_thermal_mult, self.min_mass),
            self.mass, self.den, self.rad, z if atm_pres0 is None else atm_pres0,
            z if atm_scale_h is None else atm_scale_h, z if atm_den0 is None else atm_den0,
            z if atm_h is None else atm_h,
            np.zeros((n, 3), dtype=np.int64) if base_color is None else base_color)
        out["stellar_class"] = STELLAR_CLASSES[out["stellar_class"]]
        return out

    # phys_editor.py:547-551
    def check_collision(self):
        return self.dev.check_collision()

    # phys_editor.py:381-384
    def kepler_basic(self):
        if self._small_eager():
            rec = np.empty(4 + 14 * len(self.mass))
            order = None if self._order_on_device else self._order()
            self.dev.editor_ticks(self.gc, self.coi_coef, order, 0, do_kepler_basic=True, frame_out=rec)
            self._order_on_device = True
            self._unpack_frame(rec, ("focus", "ecc_v", "semi_major", "semi_minor", "periapsis_arg", "coi"))
            return
        self.dev.kepler_basic(self.gc, self.coi_coef)
        self._touched("focus", "ecc_v", "semi_major", "semi_minor", "periapsis_arg", "coi")

    # phys_editor.py:519-544  -> ndarray (2, N, P)
    def curve(self):
        return self.dev.editor_curve(self.curve_points, self.ell_t, self.par_t)

    # phys_editor.py:554-565
    def inelastic_collision(self):
        body_del, body_add = self.check_collision()
        if body_del is not None:
            self._merge(body_del, body_add)
        return body_del

    def _merge(self, body_del, body_add):
        """:556-564 for an already detected pair."""
        mass_r = self.mass[body_del] + self.mass[body_add]
        mom1 = self._row("vel", body_add) * self.mass[body_add]
        mom2 = self._row("vel", body_del) * self.mass[body_del]
        velr = (mom1 + mom2) / mass_r
        parent = int(self._row("parents", body_add))
        self._set_row("rel_vel", body_add, velr - self._row("rel_vel", parent))
        self.set_body_mass(body_add, mass_r)
        self.del_body(body_del)

    def _order(self):
        """argsort(mass)[-1::-1], recomputed only when the masses changed."""
        if self._order_cache is None or len(self._order_cache) != len(self.mass):
            self._order_cache = mass_order(self.mass)
        return self._order_cache

    def _mass_changed(self):
        self._order_cache = None
        self._order_on_device = False
        self._rad_dirty = True

    # phys_editor.py:844-857
    def set_body_mass(self, body, mass):
        self._set_row("mass", body, mass)
        self._mass_changed()
        old_largest = int(self.largest)
        self.find_parents()
        self.simplified_orbit_coi()
        if self.largest != old_largest:
            self._reroot(old_largest)

    def _reroot(self, old_largest):
        """:848-857 / :279-289: new root to the origin, first row."""
        self.dev.translate_to(self.largest)
        self._touched("pos")
        if old_largest is not None:
            self._set_row("rel_vel", old_largest, self._row("rel_vel", self.largest) * -1)
        self._set_row("vel", self.largest, np.zeros(2))
        if self.largest != 0:
            self.set_root(self.largest)

    # phys_editor.py:293-313
    def set_root(self, body):
        self.dev.swap_rows(0, body)
        for name in ("mass", "den", "rad"):
            arr = getattr(self, name)
            arr[0], arr[body] = arr[body], arr[0]
        self._mass_changed()
        if len(self.names) > body:
            self.names[0], self.names[body] = self.names[body], self.names[0]
        for name in _AUX_SWAPPED:
            arr = getattr(self, name)
            if len(arr) > body:
                if arr.ndim == 1:
                    arr[0], arr[body] = arr[body], arr[0]
                else:
                    arr[[0, body]] = arr[[body, 0]]
        self._touched("pos", "vel", "rel_vel", "parents")
        self.simplified_orbit_coi()
        self.find_parents()
        self.largest = 0

    # phys_editor.py:250-290
    def del_body(self, delete):
        if len(self.mass) > 1:
            self.pull()                      # bring every mirror up to date, then drop the row
            self.dev.delete_row(delete)
            self._delete_host_row(delete)
            self._mass_changed()
            if len(self.names) > delete:
                self.names = np.delete(self.names, delete)
            for name in _AUX_DELETED:
                arr = getattr(self, name)
                if len(arr) > delete:
                    setattr(self, name, np.delete(arr, delete, axis=0))
            self.simplified_orbit_coi()
            self.find_parents()
            if self.largest == delete:
                self.find_parents()
                self.simplified_orbit_coi()
                self.dev.translate_to(self.largest)
                self._touched("pos")
                self._set_row("vel", self.largest, np.zeros(2))
                if self.largest != 0:
                    self.set_root(self.largest)
            return 0
        return 1

    # phys_editor.py:204-247
    def add_body(self, data):
        self.pull()
        old_largest = int(self.largest)
        n = len(self.mass) + 1
        old = {name: getattr(self, name) for name in _FIELDS}
        self._alloc_host(n)
        for name, arr in old.items():
            getattr(self, name)[:n - 1] = arr
        self.names = np.append(self.names, data.get("name", "body%d" % (n - 1)))
        self.temp = np.append(self.temp, 0)
        self.luminosity = np.append(self.luminosity, 0)
        self.stellar_class = np.append(self.stellar_class, "Unknown")
        self.color = np.vstack((self.color, (0, 0, 0)))
        self.base_color = np.vstack((self.base_color, data.get("color", (0, 0, 0))))
        self.atm_h = np.append(self.atm_h, 0)
        self.types = np.append(self.types, 0)
        self.rad_sc = np.append(self.rad_sc, 0)
        self.atm_pres0 = np.append(self.atm_pres0, data.get("atm_pres0", 0.0))
        self.atm_scale_h = np.append(self.atm_scale_h, data.get("atm_scale_h", 0.0))
        self.atm_den0 = np.append(self.atm_den0, data.get("atm_den0", 0.0))
        self.mass[-1] = data["mass"]
        self.den[-1] = data["density"]
        self.rad[:] = body_radius(self.mass, self.den, self.rad_mult)
        self.pos[-1] = data["position"]
        self._mass_changed()
        self.dev.bodies_resize(n)
        self._push()
        self.simplified_orbit_coi()
        self.find_parents()
        self._set_row("rel_vel", n - 1, np.asarray(data["velocity"], dtype=np.float64))
        if data["mass"] > self.mass[old_largest]:
            self.largest = n - 1
            self._reroot(old_largest)
        self.pull()
        for body in mass_order(self.mass):       # :245-247
            self.vel[body] = self.rel_vel[body] + self.vel[self.parents[body]]
        self._push("vel")

    # phys_editor.py:860-880 (mutators used by the editor UI)
    def set_body_den(self, body, density):
        self._set_row("den", body, max(density, 0.05))
        self._rad_dirty = True
        self.simplified_orbit_coi()

    def set_body_pos(self, body, position):
        self._set_row("pos", body, np.asarray(position, dtype=np.float64))
        self.simplified_orbit_coi()

    def set_body_vel(self, body, velocity):
        parent = int(self._row("parents", body))
        self._set_row("rel_vel", body,
                      np.asarray(velocity, dtype=np.float64) - self._row("vel", parent))
        self.simplified_orbit_coi()

    # phys_editor.py:316-322
    def move_parent(self, body, position):
        """Move a body together with the bodies that orbit it."""
        self.pull("pos", "parents")
        shift = np.asarray(position, dtype=np.float64) - self.pos[body]
        rows = np.flatnonzero(self.parents == body)    # the root is its own parent: moved twice, as there
        self.pos[body] = position
        self.pos[rows] += shift
        self._push("pos")

    # phys_editor.py:453-515: the orbit typed into the editor's side panel -> position and
    # relative velocity of `body` (ellipses only, as the reference limits it)
    def kepler_inverse(self, body, ecc, omega_deg, pe_d, mean_anomaly_deg, true_anomaly_deg, ap_d,
                       direction):
        self.pull("pos", "vel", "rel_vel", "parents")
        parent = int(self.parents[body])
        u = self.gc * self.mass[parent]
        omega = omega_deg * np.pi / 180
        mean_anomaly = mean_anomaly_deg * np.pi / 180
        if direction > 0:
            mean_anomaly = -mean_anomaly
        ecc = 0.00001 if ecc == 0 else (0.95 if ecc >= 1 else ecc)
        if ap_d:
            a = (pe_d + ap_d) / 2
            ecc = (ap_d / a) - 1
        else:
            a = - pe_d / (ecc - 1)
        b = a * math.sqrt(1 - ecc**2)
        f = math.sqrt(a**2 - b**2)
        if true_anomaly_deg:
            cos_ta = math.cos(true_anomaly_deg * np.pi / 180)
            ea = math.acos((ecc + cos_ta) / (1 + (ecc * cos_ta)))
        else:
            ea = float(self.dev.host_solve_kepler_ell([ecc], [mean_anomaly], 1e-10)[0])
        rot = omega - np.pi
        qx, qy = a * math.cos(ea) - f, b * math.sin(ea)
        pr = np.array([qx * math.cos(rot) - qy * math.sin(rot), qx * math.sin(rot) + qy * math.cos(rot)])
        angle = _ellipse_velocity_angle(pr, a, b, f, omega)
        dist = math.sqrt(pr[0] * pr[0] + pr[1] * pr[1])
        speed = direction * math.sqrt((2 * a * u - dist * u) / (a * dist))
        self.move_parent(body, self.pos[parent] + pr)
        self.rel_vel[body] = (speed * math.cos(angle), speed * math.sin(angle))
        # absolute velocities again, heaviest first, so that the edit shows at once (:512-514)
        for i in self._order():
            self.vel[i] = self.rel_vel[i] + self.vel[self.parents[i]]
        self._push("rel_vel", "vel")

    # phys_editor.py:819-836
    def get_bodies(self):
        self.pull()
        return (self.names, self.types, self.mass, self.den, self.temp, self.luminosity,
                self.stellar_class, self.pos, self.vel, self.color, self.rad, self.rad_sc,
                self.surf_grav)

    def get_atmosphere(self):
        return self.atm_pres0, self.atm_scale_h, self.atm_den0, self.atm_h

    def get_base_color(self):
        return self.base_color

    def get_body_orbits(self):
        self.pull()
        return self.semi_major, self.semi_minor, self.coi, self.parents

    # phys_editor.py:839-841, 883-897
    def set_body_name(self, body, name):
        self.names[body] = name

    def set_body_base_color(self, body, color):
        self.base_color[body] = color
        self._rad_dirty = True

    def set_body_atmos(self, body, atm_pres0, atm_scale_h, atm_den0):
        pick = lambda v: v[body] if isinstance(v, (list, np.ndarray)) else v
        self.atm_pres0[body] = pick(atm_pres0)
        self.atm_scale_h[body] = pick(atm_scale_h)
        self.atm_den0[body] = pick(atm_den0)
        self._rad_dirty = True

    def frame_calls(self, warp=1):
        """One editor frame as the individual method calls of editor.py:1503-1523."""
        deleted = []
        for _ in range(warp):
            self.gravity()
            self.body()
            d = self.inelastic_collision()
            if d is not None:
                deleted.append(int(d))
        self.kepler_basic()
        return deleted

    _FRAME_FIELDS = ("pos", "vel", "rel_vel", "parents", "focus", "ecc_v", "semi_major",
                     "semi_minor", "periapsis_arg", "coi")

    def frame(self, warp=1):
        """The same frame through the fused driver (vsb_editor_ticks): the warp loop, the
        collision checks and kepler_basic run on the device without host round trips (one
        kernel launch and one 32-byte read-back for up to 1024 bodies); the host only steps
        in to apply a merge.  Bit-identical to frame_calls().  Returns the deleted indices."""
        deleted = []
        left = int(warp)
        eager = self.sync == "eager"
        while True:
            self.body()          # editor.py:1505 -- rad only changes after a merge
            n = len(self.mass)
            rec = np.empty(4 + 14 * n) if eager else None
            order = None if self._order_on_device else self._order()
            done, body_del, body_add = self.dev.editor_ticks(
                self.gc, self.coi_coef, order, left, do_kepler_basic=True, frame_out=rec)
            self._order_on_device = True
            left -= done
            if body_del is None:
                if eager:
                    self._unpack_frame(rec)
                else:
                    self._stale.update(self._FRAME_FIELDS)
                break
            self._stale.update(("parents", "rel_vel", "vel", "pos"))
            self._merge(body_del, body_add)
            deleted.append(int(body_del))
            if left == 0:
                self.kepler_basic()
                if eager:
                    self.pull()
                break
        return deleted

    def _unpack_frame(self, rec, fields=None):
        """Frame record of vsb_editor_ticks -> host mirrors (in place); `fields`: only those."""
        n = len(self.mass)
        off = 4
        for k in self._FRAME_FIELDS:
            w = 2 if k in _WIDE else 1
            chunk = rec[off:off + n * w]
            off += n * w
            if fields is not None and k not in fields:
                continue
            dst = getattr(self, k)
            if k == "parents":
                dst[:] = chunk.view(np.int64)
            else:
                dst[...] = chunk.reshape(dst.shape)
            self._stale.discard(k)


class AllPairsPhysics(_DeviceBodies):
    """MODE_ALLPAIRS: every body attracts every other body
    (documentation/orbital-physics/01-complex-orbit-model.txt:4-12), integrator
    v += a; P += v, collisions as check_collision (phys_editor.py:547-551) with the
    documented inelastic merge (documentation/orbital-physics/09-collisions.txt:3-6).
    Same call sequence as the editor: gravity(); body(); inelastic_collision()."""

    def __init__(self, device=None, sync="lazy"):
        super().__init__(device, sync)
        self.gc, self.rad_mult = 1.0, 179.0
        self._rad_rows = None           # rows whose radius is out of date (None: all of them)
        self._alloc_host(0)

    def load_system(self, conf, body_data, body_orb_data):
        self.gc, self.rad_mult = conf["gc"], conf["rad_mult"]
        mass = np.array(body_data["mass"], dtype=np.float64)
        n = len(mass)
        self._alloc_host(n)
        self.mass[:] = mass
        self.den[:] = body_data["den"]
        self.rad[:] = body_radius(self.mass, self.den, self.rad_mult)
        self.pos[:] = body_orb_data["pos"]
        self.vel[:] = body_orb_data["vel"]
        self._stale.clear()
        self._rad_rows = set()
        self.dev.bodies_resize(n)
        self._push("mass", "den", "rad", "pos", "vel")

    def gravity(self, nsteps=1):
        self.dev.allpairs_step(self.gc, nsteps)
        self._touched("pos", "vel")

    def acceleration(self):
        self.dev.allpairs_accel(self.gc)
        return self.dev.get(F_ACC)

    def body(self):
        """rad = rad_mult * cbrt(3 (mass / den) / (4 pi)), phys_editor.py:577-581, host numpy (it
        feeds the bit-exact collision predicate).  Only rows whose mass changed since the last
        call are re-evaluated: np.cbrt is elementwise, a one-row call gives the bits of the
        full-array call (3.7 ms per tick saved at 262 144 bodies)."""
        rows = self._rad_rows
        if rows is None:
            rad = body_radius(self.mass, self.den, self.rad_mult)
            if not np.array_equal(rad, self.rad):
                self.rad[:] = rad
                self._push("rad")
        else:
            for i in sorted(rows):
                r = body_radius(self.mass[i:i + 1], self.den[i:i + 1], self.rad_mult)[0]
                if r != self.rad[i]:
                    self.rad[i] = r
                    self.dev.set(F_RAD, self.rad[i:i + 1], first=int(i))
        self._rad_rows = set()

    def check_collision(self):
        return self.dev.check_collision()

    def inelastic_collision(self):
        body_del, body_add = self.check_collision()
        if body_del is None:
            return None
        mass_r = self.mass[body_del] + self.mass[body_add]
        velr = (self._row("vel", body_add) * self.mass[body_add]
                + self._row("vel", body_del) * self.mass[body_del]) / mass_r
        self._set_row("vel", body_add, velr)
        self._set_row("mass", body_add, mass_r)
        self.dev.delete_row(body_del)
        stale = set(self._stale)
        self._delete_host_row(body_del)
        self._stale = stale
        if self._rad_rows is not None:      # the survivor's radius is due at the next body()
            moved = {j - (j > body_del) for j in self._rad_rows if j != body_del}
            moved.add(body_add - (body_add > body_del))
            self._rad_rows = moved
        return body_del, body_add

    _LIVE = ("mass", "den", "rad", "pos", "vel")     # the fields this model uses

    def _delete_host_row(self, i):
        """np.delete of row i in every host mirror (phys_editor.py:253-278) without re-allocating
        them per merge: live mirrors shift their tail down in place, mirrors that are stale (the
        device copy is authoritative) or unused by this model (all zeros) just lose their last
        row.  Mirrors are re-bound like in the reference (aliases held by a caller go stale)."""
        for name in _FIELDS:
            arr = getattr(self, name)
            if name in self._LIVE and name not in self._stale:
                arr[i:-1] = arr[i + 1:]
            setattr(self, name, arr[:-1])

    def tick(self):
        self.gravity()
        self.body()
        return self.inelastic_collision()
