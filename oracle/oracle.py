"""Python face of the CPU parity oracle (TEST INFRASTRUCTURE ONLY).

Loads oracle/libvs_oracle.so (built from vs_oracle.c by `make -C oracle`) and
adds numpy-level restatements of the reference's O(N) host logic (merge
bookkeeping, root swap) that sits between the O(N^2) loops.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package
volatilespace_b200 never does.

Pinning status: pinned against outputs of the live reference (numba path)
committed under tests/golden/ by oracle/gen_golden.py -- the reference has no
golden vectors or tests of its own.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvs_oracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"],
                          stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    i64, f64 = C.c_int64, C.c_double
    _u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
    sigs = {
        "vso_num_threads": (C.c_int, []),
        "vso_set_num_threads": (None, [C.c_int]),
        "vso_set_exact_trig": (None, [C.c_int]),
        "vso_find_parents": (None, [i64, _i64p, _f64p, _f64p, _i64p]),
        "vso_gravity_one": (None, [i64, i64, _f64p, f64, f64, f64, _f64p]),
        "vso_gravity_ref": (None, [i64, _i64p, _f64p, _f64p, f64, _f64p, _f64p, _f64p, _i64p]),
        "vso_allpairs_accel": (None, [i64, _f64p, _f64p, f64, i64, i64, _f64p]),
        "vso_allpairs_step": (None, [i64, _f64p, _f64p, _f64p, f64, i64]),
        "vso_kepler_basic": (None, [i64, _i64p, _f64p, _f64p, _f64p, f64, f64,
                                    _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]),
        "vso_simplified_orbit_coi": (i64, [i64, _i64p, _f64p, _f64p, _f64p, f64, f64, _f64p]),
        "vso_check_collision": (None, [i64, _f64p, _f64p, _f64p,
                                       C.POINTER(i64), C.POINTER(i64)]),
        "vso_solve_kepler_ell": (f64, [f64, f64, f64]),
        "vso_newton_root_kepler_hyp": (f64, [f64, f64, f64]),
        "vso_newton_root_kepler_ell": (f64, [f64, f64, f64]),
        "vso_solve_kepler_ell_batch": (None, [i64, _f64p, _f64p, f64, _f64p]),
        "vso_newton_root_kepler_hyp_batch": (None, [i64, _f64p, _f64p, _f64p, _f64p]),
        "vso_calc_orb_one": (None, [C.c_int, i64, i64, _f64p, f64, f64, f64, f64, f64, _f64p]),
        "vso_body_move": (None, [i64, f64, _i64p, _i64p] + [_f64p] * 10),
        "vso_vessel_move": (None, [i64, f64, _i64p, _f64p] + [_f64p] * 10),
        "vso_curve_points": (None, [f64, f64, f64, f64, _f64p, i64, _f64p]),
        "vso_curves_all": (None, [i64, _f64p, _f64p, _f64p, _f64p, _f64p, i64, _f64p]),
        "vso_curve_move_to": (None, [i64, i64, _f64p, _f64p, _i64p, _f64p, _f64p, _f64p]),
        "vso_to_kepler": (None, [i64, _i64p, _f64p, _f64p, _f64p, f64, f64, _i64p, _f64p, _f64p,
                                 _f64p, _f64p, _f64p]),
        "vso_to_newton": (C.c_int, [i64, _f64p, _f64p, _f64p, _f64p, _f64p, _i64p, _f64p, f64,
                                    _f64p, _f64p]),
        "vso_solve_quartic": (None, [f64, f64, f64, f64, f64, _f64p]),
        "vso_ell_hyp_intersect_circle": (C.c_int, [f64, f64, f64, f64, f64, f64, _f64p]),
        "vso_predict_enter_coi": (None, [_f64p, _f64p, f64, C.c_int, _f64p]),
        "vso_point_between": (C.c_int, [f64, f64, f64, f64]),
        "vso_check_warp": (None, [i64] + [_f64p] * 9 + [f64, _u8p, _i64p]),
        "vso_enter_coi_service": (None, [i64, _f64p, _f64p, _f64p, _f64p, C.c_int, _u8p]),
        "vso_cross_scan": (None, [i64] + [_f64p] * 7 + [_i64p]),
        "vso_kepler_to_velocity": (None, [_f64p, f64, f64, f64, f64, f64, _f64p]),
        "vso_velocity_to_kepler": (None, [_f64p, _f64p, f64, C.c_int, _f64p]),
        "vso_cross_apply": (C.c_int, [C.c_int, _f64p, i64, i64, _f64p, f64, _f64p]),
        "vso_vessel_points": (None, [i64, _i64p, i64, _f64p, _i64p, _f64p, _i64p, i64, i64, i64,
                                     C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p]),
        "vso_editor_curve": (None, [i64, i64, _f64p, _f64p, _f64p, _f64p, _f64p, _i64p,
                                    _f64p, _f64p, _f64p, _f64p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def mass_order(mass):
    """argsort(mass)[-1::-1] exactly as phys_editor.py:353,375 computes it."""
    return _i(np.argsort(mass)[-1::-1])


# ---------------------------------------------------------------- free functions
def find_parents(mass, pos, coi, order=None):
    n = len(mass)
    order = mass_order(mass) if order is None else _i(order)
    out = np.zeros(n, dtype=np.int64)
    lib().vso_find_parents(n, order, _f(pos), _f(coi), out)
    return out


def gravity_one(body, parent, mass, rel_pos, gc):
    out = np.zeros(2)
    lib().vso_gravity_one(int(body), int(parent), _f(mass), float(rel_pos[0]),
                          float(rel_pos[1]), float(gc), out)
    return out


def allpairs_accel(mass, pos, gc, row0=0, row1=None):
    n = len(mass)
    row1 = n if row1 is None else row1
    acc = np.zeros((row1 - row0, 2))
    lib().vso_allpairs_accel(n, _f(mass), _f(pos), float(gc), row0, row1, acc)
    return acc


def allpairs_step(mass, pos, vel, gc, nsteps=1):
    """Returns new (pos, vel) after nsteps ticks of v += a; P += v."""
    pos = _f(pos).copy()
    vel = _f(vel).copy()
    lib().vso_allpairs_step(len(mass), _f(mass), pos, vel, float(gc), int(nsteps))
    return pos, vel


def check_collision(mass, pos, rad):
    d, a = C.c_int64(-1), C.c_int64(-1)
    lib().vso_check_collision(len(mass), _f(mass), _f(pos), _f(rad), C.byref(d), C.byref(a))
    if d.value < 0:
        return None, None
    return int(d.value), int(a.value)


def kepler_basic(parents, mass, pos, vel, gc, coi_coef):
    n = len(mass)
    focus, sma, smi, pea, coi = (np.zeros(n) for _ in range(5))
    ecc_v = np.zeros((n, 2))
    lib().vso_kepler_basic(n, _i(parents), _f(mass), _f(pos), _f(vel), float(gc),
                           float(coi_coef), focus, ecc_v, sma, smi, pea, coi)
    return focus, ecc_v, sma, smi, pea, coi


def simplified_orbit_coi(mass, pos, vel, gc, coi_coef, order=None):
    n = len(mass)
    order = mass_order(mass) if order is None else _i(order)
    coi = np.zeros(n)
    largest = lib().vso_simplified_orbit_coi(n, order, _f(mass), _f(pos), _f(vel),
                                             float(gc), float(coi_coef), coi)
    return int(largest), coi


def solve_kepler_ell(ecc, ma, tol=1e-10):
    if np.ndim(ecc) == 0 and np.ndim(ma) == 0:
        return lib().vso_solve_kepler_ell(float(ecc), float(ma), float(tol))
    ecc, ma = np.broadcast_arrays(_f(ecc), _f(ma))
    ecc, ma = _f(ecc), _f(ma)
    out = np.zeros(ecc.shape)
    lib().vso_solve_kepler_ell_batch(ecc.size, ecc.ravel(), ma.ravel(), float(tol), out.ravel())
    return out


def newton_root_kepler_hyp(ecc, ma, guess):
    if np.ndim(ecc) == 0 and np.ndim(ma) == 0:
        return lib().vso_newton_root_kepler_hyp(float(ecc), float(ma), float(guess))
    ecc, ma, guess = (_f(x) for x in np.broadcast_arrays(_f(ecc), _f(ma), _f(guess)))
    out = np.zeros(ecc.shape)
    lib().vso_newton_root_kepler_hyp_batch(ecc.size, ecc.ravel(), ma.ravel(), guess.ravel(),
                                           out.ravel())
    return out


def calc_orb(kind, ref, mass_ref_table, body_mass, gc, coi_coef, a, ecc):
    """Vectorised C8.  kind='body' (phys_body.py:27-60) or 'vessel'
    (phys_vessel.py:82-110).  Returns dict of arrays b,f,coi,pe_d,ap_d,period,n,u."""
    n = len(a)
    out = np.zeros((n, 8))
    tmp = np.zeros(8)
    mt = _f(mass_ref_table)
    with_coi = 1 if kind == "body" else 0
    for i in range(n):
        bm = float(body_mass[i]) if with_coi else 0.0
        lib().vso_calc_orb_one(with_coi, i, int(ref[i]), mt, bm, float(gc), float(coi_coef),
                               float(a[i]), float(ecc[i]), tmp)
        out[i] = tmp
    keys = ["b", "f", "coi", "pe_d", "ap_d", "period", "n", "u"]
    return {k: np.ascontiguousarray(out[:, j]) for j, k in enumerate(keys)}


def body_move(warp, ref, coi, a, b, f, ecc, pea, n, dr, ma, ea, pos):
    """C1; returns new (pos, ma, ea)."""
    order = _i(np.argsort(coi)[-1::-1])
    ma, ea, pos = _f(ma).copy(), _f(ea).copy(), _f(pos).copy()
    lib().vso_body_move(len(a), float(warp), order, _i(ref), _f(a), _f(b), _f(f), _f(ecc),
                        _f(pea), _f(n), _f(dr), ma, ea, pos)
    return pos, ma, ea


def vessel_move(warp, ref, body_pos, a, b, f, ecc, pea, n, dr, ma, ea):
    """C4; returns new (pos, ma, ea)."""
    ma, ea = _f(ma).copy(), _f(ea).copy()
    pos = np.zeros((len(a), 2))
    lib().vso_vessel_move(len(a), float(warp), _i(ref), _f(body_pos), _f(a), _f(b), _f(f),
                          _f(ecc), _f(pea), _f(n), _f(dr), ma, ea, pos)
    return pos, ma, ea


def curve_points(ecc, a, b, pea, t):
    t = _f(t)
    out = np.zeros((len(t), 2))
    lib().vso_curve_points(float(ecc), float(a), float(b), float(pea), t, len(t), out)
    return out


def curves_all(ecc, a, b, pea, t):
    t = _f(t)
    out = np.zeros((len(a), len(t), 2))
    lib().vso_curves_all(len(a), _f(ecc), _f(a), _f(b), _f(pea), t, len(t), out)
    return out


def curve_move_to(curves, body_pos, ref, f, pea):
    curves = _f(curves)
    out = np.zeros_like(curves)
    lib().vso_curve_move_to(curves.shape[0], curves.shape[1], curves, _f(body_pos), _i(ref),
                            _f(f), _f(pea), out)
    return out


def editor_curve(ecc_v, semi_major, semi_minor, focus, periapsis_arg, parents, pos, P):
    n = len(semi_major)
    ell_t = np.linspace(-np.pi, np.pi, P)
    par_t = np.linspace(-np.pi - 1, np.pi + 1, P)
    out = np.zeros((2, n, P))
    lib().vso_editor_curve(n, P, _f(ecc_v), _f(semi_major), _f(semi_minor), _f(focus),
                           _f(periapsis_arg), _i(parents), _f(pos), ell_t, par_t, out)
    return out


def culling(coords, radius, screen_bounds, zoom):
    """phys_shared.culling, phys_shared.py:224-231 (numpy restatement, same expressions)."""
    coords, radius, sb = _f(coords), _f(radius), _f(screen_bounds).reshape(2, 2)
    min_dim = (coords.T + radius).T + 2 / zoom > np.array([sb[0, 0], sb[1, 1]])
    max_dim = (coords.T - radius).T + 2 / zoom < np.array([sb[1, 0], sb[0, 1]])
    return np.logical_and(np.logical_and(min_dim[:, 0], min_dim[:, 1]),
                          np.logical_and(max_dim[:, 0], max_dim[:, 1]))


def body_culling(pos, radius, atm_h, coi, ref, a, sim_screen, zoom, screen_x):
    """phys_body.Physics.culling, phys_body.py:201-218."""
    n = len(coi)
    visible_bodies = np.arange(n)[culling(pos, (radius + atm_h) * zoom, sim_screen, zoom)]
    coi_truth = culling(pos, coi, sim_screen, zoom)
    visible_coi = np.concatenate(([0], np.arange(n)[coi_truth]))
    t = np.logical_and(coi_truth[ref], coi[ref] > 8 / zoom)
    t = np.logical_or(t, ref == 0)
    t = np.logical_and(t, a * zoom < screen_x * 15)
    return visible_bodies, visible_coi, np.arange(n)[t]


def vessel_culling(pos, ref, pe_d, body_coi, visible_coi, sim_screen, zoom, screen_x):
    """phys_vessel.Physics.culling, phys_vessel.py:258-273."""
    n = len(ref)
    visible_vessels = np.arange(n)[culling(pos, np.array([10] * n) / zoom, sim_screen, zoom)]
    coi_truth = np.array([False] * len(body_coi))
    coi_truth[visible_coi] = True
    t = np.logical_and(coi_truth[ref], body_coi[ref] > 8 / zoom)
    t = np.logical_or(t, ref == 0)
    t = np.logical_and(t, pe_d * zoom < screen_x * 15)
    return visible_vessels, np.arange(n)[t]


def solve_quartic(a, b, c, d, e):
    """quartic_solver.solve_quartic, quartic_solver.py:57-93 -> 4 complex roots."""
    r = np.zeros(8)
    lib().vso_solve_quartic(float(a), float(b), float(c), float(d), float(e), r)
    return r[0::2] + 1j * r[1::2]


def ell_hyp_intersect_circle(a, b, ecc, c_x, c_y, r):
    """orbit_intersect.py:17-52 -> array of eccentric anomalies ([nan] when there is none)."""
    ea = np.full(4, np.nan)
    cnt = lib().vso_ell_hyp_intersect_circle(float(a), float(b), float(ecc), float(c_x),
                                             float(c_y), float(r), ea)
    return ea[:cnt].copy() if cnt else np.array([np.nan])


def predict_enter_coi(vessel_data, body_data, tol, steps_limit):
    """orbit_intersect.py:125-174 -> (ea, b_ea, ma, b_ma, t)."""
    out = np.zeros(5)
    lib().vso_predict_enter_coi(_f(vessel_data), _f(body_data), float(tol), int(steps_limit), out)
    return tuple(out)


def vessel_points(vessel12, v_ref, body12, body_ref, sel=None, entered_coi=None, left_coi=None,
                  left_coi_prev_ref=None, only_enter_coi=False, steps_limit=300, coi_leave=None):
    """phys_vessel.Physics.points (phys_vessel.py:372-474) for the vessels in `sel` (all when
    None) -> body_impact (nv,2), body_enter_atm (nv,2), coi_leave (nv,2), coi_enter (nv,6);
    rows of vessels outside `sel` are NaN (coi_leave: the array passed in)."""
    vessel12, body12 = _f(vessel12), _f(body12)
    nv = len(vessel12)
    sel = np.arange(nv, dtype=np.int64) if sel is None else _i(sel)
    none = lambda x: -1 if x is None else int(x)
    bi, ba = np.full((nv, 2), np.nan), np.full((nv, 2), np.nan)
    cl = np.full((nv, 2), np.nan) if coi_leave is None else _f(coi_leave).copy()
    ce = np.full((nv, 6), np.nan)
    lib().vso_vessel_points(len(sel), sel, len(body12), vessel12, _i(v_ref), body12, _i(body_ref),
                            none(entered_coi), none(left_coi), none(left_coi_prev_ref),
                            int(bool(only_enter_coi)), int(steps_limit), bi, ba, cl, ce)
    return bi, ba, cl, ce


def check_warp(ecc, dr, ea, ma, n, body_impact, body_enter_atm, coi_leave, coi_enter, warp, hold):
    """phys_vessel.Physics.check_warp (phys_vessel.py:506-560) -> (situation, alarm_vessel,
    len(physical_hold)) with None as in the reference, and the new hold mask."""
    hold = np.ascontiguousarray(hold, dtype=np.uint8).copy()
    out = np.zeros(3, dtype=np.int64)
    lib().vso_check_warp(len(ecc), _f(ecc), _f(dr), _f(ea), _f(ma), _f(n), _f(body_impact),
                         _f(body_enter_atm), _f(coi_leave), _f(coi_enter), float(warp), hold, out)
    opt = lambda x: None if x < 0 else int(x)
    return (opt(out[0]), opt(out[1]), int(out[2])), hold


def enter_coi_service(dr, prev_ea, ea, coi_enter, entered_coi=None, left_coi=None):
    """predict_enter_coi_service (phys_vessel.py:744-758): the vessels it calls points(v, True) on."""
    flags = np.zeros(len(dr), dtype=np.uint8)
    lib().vso_enter_coi_service(len(dr), _f(dr), _f(prev_ea), _f(ea), _f(coi_enter),
                                int(entered_coi is not None or left_coi is not None), flags)
    return np.flatnonzero(flags)


def cross_scan(ecc, dr, prev_ea, ea, body_impact, coi_leave, coi_enter):
    """cross_coi (phys_vessel.py:563-632), detection: (vessel, kind) of the first crossing, kind
    1 = enter, 2 = leave; (None, 0) when nothing crosses."""
    out = np.zeros(2, dtype=np.int64)
    lib().vso_cross_scan(len(ecc), _f(ecc), _f(dr), _f(prev_ea), _f(ea), _f(body_impact),
                         _f(coi_leave), _f(coi_enter), out)
    return (None if out[0] < 0 else int(out[0])), int(out[1])


def cross_apply(kind, vessel8, ref, new_ref, body7, gc):
    """cross_coi, the orbit change (phys_vessel.py:585-629) -> (a, ecc, pea, ma, dr, new_ref) or
    None where the reference returns None."""
    out = np.zeros(6)
    rc = lib().vso_cross_apply(int(kind), _f(vessel8), int(ref), int(new_ref), _f(body7), float(gc),
                               out)
    return None if rc else out


def kepler_to_velocity(rel_pos, a, ecc, pe_arg, u, dr):
    out = np.zeros(2)
    lib().vso_kepler_to_velocity(_f(rel_pos), float(a), float(ecc), float(pe_arg), float(u),
                                 float(dr), out)
    return out


def velocity_to_kepler(rel_pos, rel_vel, u, failsafe=0):
    out = np.zeros(5)
    lib().vso_velocity_to_kepler(_f(rel_pos), _f(rel_vel), float(u), int(failsafe), out)
    return out


def to_kepler(mass, pos, vel, gc, coi_coef):
    """N1: convert.to_kepler, convert.py:193-261 -> dict like the reference's return value."""
    n = len(mass)
    ref = np.zeros(n, dtype=np.int64)
    a, ecc, pe, ma, dr = (np.zeros(n) for _ in range(5))
    lib().vso_to_kepler(n, mass_order(mass), _f(mass), _f(pos), _f(vel), float(gc),
                        float(coi_coef), ref, a, ecc, pe, ma, dr)
    return {"kepler": True, "a": a, "ecc": ecc, "pe_arg": pe, "ma": ma, "ref": ref, "dir": dr}


def to_newton(mass, orb, gc, coi_coef):
    """N1: convert.to_newton, convert.py:132-190 (ValueError for hyperbolic bodies, as there)."""
    n = len(mass)
    pos, vel = np.zeros((n, 2)), np.zeros((n, 2))
    rc = lib().vso_to_newton(n, _f(mass), _f(orb["a"]), _f(orb["ecc"]), _f(orb["pe_arg"]),
                             _f(orb["ma"]), _i(orb["ref"]), _f(orb["dir"]), float(gc), pos, vel)
    if rc != 0:
        raise ValueError("math domain error")
    return {"kepler": False, "pos": pos, "vel": vel}


# phys_shared.py:13-17
GC_REAL = 6.674 * 10**-11
C_LIGHT = 299792458
SIGMA = 5.670374419 * 10**-8
LS = 3.828 * 10**26
MS = 1.9884 * 10**30


def body_thermal(mass, den, rad, atm_pres0, atm_scale_h, atm_den0, atm_h, base_color, gc,
                 rad_mult, mass_thermal_mult, min_mass):
    """N4: Physics.body (phys_editor.py:583-599), temp_color (:602-651), classify (:654-670),
    numpy restatement.  Returns dict."""
    import math
    mass, den, rad = _f(mass), _f(den), _f(rad)
    atm_h = _f(atm_h).copy()
    surf_grav = gc * mass / rad**2
    for b in range(len(mass)):
        if atm_pres0[b] != 0 and atm_scale_h[b] != 0 and atm_den0[b] != 0:
            atm_h[b] = -atm_scale_h[b] * math.log(0.001 / atm_den0[b]) * rad_mult
    tm = mass * mass_thermal_mult
    tr = np.cbrt(3 * (tm / den) / (4 * np.pi))
    lum = LS * (tm / MS)**3.5
    temp = lum**(1 / 4) / (np.sqrt(tr) * np.sqrt(2) * np.pi**(1 / 4) * SIGMA**(1 / 4))
    rad_sc = 2 * tm * GC_REAL / C_LIGHT**2
    rad_sc = rad_sc * rad / tr
    atm_h = np.where(rad_sc > rad, 0, atm_h)
    base = np.asarray(base_color)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        fg = np.where((temp > 1000) & (temp <= 2300), 1 - 1 / (1300 / (temp - 1000)), -1)
        color = np.copy(base)
        for ch, hot in ((0, 255), (1, 184), (2, 111)):
            color[:, ch] = np.where(fg != -1, base[:, ch] * fg + hot * (1 - fg), base[:, ch])
        color[:, 0] = np.where(temp > 2300, 130 + 125 / (1 + (temp / 5922)**55.525)**0.065,
                               color[:, 0])
        color[:, 1] = np.where((temp > 2300) & (temp <= 6000),
                               257 - 73 / (1 + (temp / 4742)**6.687), color[:, 1])
        color[:, 1] = np.where(temp > 6000, 170 + 165766330 / (1 + (temp / 24)**2.651),
                               color[:, 1])
        color[:, 2] = np.where(temp > 2300, 283 - 176 / (1 + (temp / 4493)**5.792), color[:, 2])
    color[rad_sc > rad] = [0, 0, 0]
    color = np.clip(color.astype(int), 0, 255)
    types = np.zeros(len(mass))
    types = np.where(mass > min_mass, 1, types)
    types = np.where(den < 1500, 2, types)
    types = np.where(temp > 1000, 3, types)
    types = np.where(rad_sc > rad, 4, types)
    sc = np.where(types == 3, "M", "Unknown")
    for thr, name in ((3900, "K"), (5300, "G"), (6000, "F"), (7300, "A"), (10000, "B"),
                      (33000, "O")):
        sc = np.where(temp > thr, name, sc)
    return dict(surf_grav=surf_grav, atm_h=atm_h, luminosity=lum, temp=temp, rad_sc=rad_sc,
                color=color, types=types, stellar_class=sc)


def body_radius(mass, den, rad_mult):
    """A5 radius only, phys_editor.py:580-581 (host numpy on both sides)."""
    volume = mass / den
    return rad_mult * np.cbrt(3 * volume / (4 * np.pi))


# ---------------------------------------------------------------- editor state
class EditorOracle:
    """Restatement of phys_editor.Physics restricted to the arrays the hot path
    reads or writes (phys_editor.py:112-149): mass, den, rad, pos, vel, rel_vel,
    coi, parents, largest and the kepler_basic outputs."""

    def __init__(self, mass, den, pos, vel, gc=1.0, rad_mult=179.0, coi_coef=0.7):
        self.gc, self.rad_mult, self.coi_coef = float(gc), float(rad_mult), float(coi_coef)
        # load_system, phys_editor.py:169-201
        self.mass = _f(mass).copy()
        self.den = _f(den).copy()
        self.rad = body_radius(self.mass, self.den, self.rad_mult)
        self.pos = _f(pos).copy()
        self.vel = _f(vel).copy()
        self.parents = np.array([], dtype=np.int64)
        self.simplified_orbit_coi()
        self.find_parents()
        self.rel_vel = self.vel - self.vel[self.parents]
        n = len(self.mass)
        self.focus = np.zeros(n)
        self.semi_major = np.zeros(n)
        self.semi_minor = np.zeros(n)
        self.periapsis_arg = np.zeros(n)
        self.ecc_v = np.zeros((n, 2))
        self.body()

    # phys_editor.py:325-346
    def simplified_orbit_coi(self):
        self.largest, self.coi = simplified_orbit_coi(self.mass, self.pos, self.vel, self.gc,
                                                      self.coi_coef)

    # phys_editor.py:350-354
    def find_parents(self):
        self.parents = find_parents(self.mass, self.pos, self.coi)

    # phys_editor.py:357-378
    def gravity(self):
        n = len(self.mass)
        order = mass_order(self.mass)
        parents = self.parents.copy() if len(self.parents) == n else np.zeros(n, np.int64)
        lib().vso_gravity_ref(n, order, self.mass, self.coi, self.gc, self.pos, self.vel,
                              self.rel_vel, parents)
        self.parents = parents

    # phys_editor.py:577-581 (radius part)
    def body(self):
        self.rad = body_radius(self.mass, self.den, self.rad_mult)

    # phys_editor.py:547-551
    def check_collision(self):
        return check_collision(self.mass, self.pos, self.rad)

    # phys_editor.py:381-384
    def kepler_basic(self):
        (self.focus, self.ecc_v, self.semi_major, self.semi_minor, self.periapsis_arg,
         self.coi) = kepler_basic(self.parents, self.mass, self.pos, self.vel, self.gc,
                                  self.coi_coef)

    # phys_editor.py:519-544
    def curve(self, P):
        return editor_curve(self.ecc_v, self.semi_major, self.semi_minor, self.focus,
                            self.periapsis_arg, self.parents, self.pos, P)

    # phys_editor.py:554-565
    def inelastic_collision(self):
        body_del, body_add = self.check_collision()
        if body_del is not None:
            mass_r = self.mass[body_del] + self.mass[body_add]
            mom1 = self.vel[body_add] * self.mass[body_add]
            mom2 = self.vel[body_del] * self.mass[body_del]
            velr = (mom1 + mom2) / mass_r
            self.rel_vel[body_add] = velr - self.rel_vel[self.parents[body_add]]
            self.set_body_mass(body_add, mass_r)
            self.del_body(body_del)
        return body_del

    # phys_editor.py:844-857
    def set_body_mass(self, body, mass):
        self.mass[body] = mass
        old_largest = int(self.largest)
        self.find_parents()
        self.simplified_orbit_coi()
        if self.largest != old_largest:
            diff = self.pos[self.largest].copy()
            self.pos -= diff
            self.rel_vel[old_largest] = self.rel_vel[self.largest] * -1
            self.vel[self.largest] = [0, 0]
            if self.largest != 0:
                self.set_root(self.largest)

    # phys_editor.py:293-313 (arrays on the hot path only; coi and the orbit
    # arrays are NOT swapped there either)
    def set_root(self, body):
        for arr in (self.mass, self.den, self.rad, self.parents):
            arr[0], arr[body] = arr[body], arr[0]
        for arr in (self.pos, self.vel, self.rel_vel):
            arr[[0, body]] = arr[[body, 0]]
        self.simplified_orbit_coi()
        self.find_parents()
        self.largest = 0

    # phys_editor.py:250-290
    def del_body(self, delete):
        if len(self.mass) > 1:
            for name in ("mass", "den", "rad", "parents"):
                setattr(self, name, np.delete(getattr(self, name), delete))
            for name in ("pos", "vel", "rel_vel"):
                setattr(self, name, np.delete(getattr(self, name), delete, axis=0))
            self.simplified_orbit_coi()
            self.find_parents()
            for name in ("focus", "semi_major", "semi_minor", "periapsis_arg"):
                setattr(self, name, np.delete(getattr(self, name), delete))
            self.ecc_v = np.delete(self.ecc_v, delete, axis=0)
            if self.largest == delete:
                self.find_parents()
                self.simplified_orbit_coi()
                diff = self.pos[self.largest].copy()
                self.pos -= diff
                self.vel[self.largest] = [0, 0]
                if self.largest != 0:
                    self.set_root(self.largest)
            return 0
        return 1

    # phys_editor.py:204-247
    def add_body(self, data):
        old_largest = int(self.largest)
        self.mass = np.append(self.mass, data["mass"])
        self.den = np.append(self.den, data["density"])
        self.rad = body_radius(self.mass, self.den, self.rad_mult)
        self.pos = np.vstack((self.pos, data["position"]))
        self.vel = np.vstack((self.vel, [0, 0]))
        self.parents = np.append(self.parents, 0)
        self.simplified_orbit_coi()
        self.find_parents()
        self.rel_vel = np.vstack((self.rel_vel, data["velocity"]))
        for name in ("focus", "semi_major", "semi_minor", "periapsis_arg"):
            setattr(self, name, np.append(getattr(self, name), 0))
        self.ecc_v = np.vstack((self.ecc_v, [0, 0]))
        if data["mass"] > self.mass[old_largest]:
            self.largest = len(self.mass) - 1
            diff = self.pos[self.largest].copy()
            self.pos -= diff
            self.rel_vel[old_largest] = self.rel_vel[self.largest] * -1
            self.vel[self.largest] = [0, 0]
            if self.largest != 0:
                self.set_root(self.largest)
        for body in mass_order(self.mass):
            self.vel[body] = self.rel_vel[body] + self.vel[self.parents[body]]

    # phys_editor.py:860-876
    def set_body_den(self, body, density):
        self.den[body] = max(density, 0.05)
        self.simplified_orbit_coi()

    def set_body_pos(self, body, position):
        self.pos[body] = position
        self.simplified_orbit_coi()

    def set_body_vel(self, body, velocity):
        self.rel_vel[body] = velocity - self.vel[self.parents[body]]
        self.simplified_orbit_coi()

    # phys_editor.py:316-322
    def move_parent(self, body, position):
        movement = position - self.pos[body]
        self.pos[body] = position
        for child in np.where(self.parents == body)[0]:
            self.pos[child] += movement

    # phys_editor.py:453-515 (orbit edited in the side panel: elements -> state vector)
    def kepler_inverse(self, body, ecc, omega_deg, pe_d, mean_anomaly_deg, true_anomaly_deg, ap_d,
                       direction):
        import math
        parent = self.parents[body]
        u = self.gc * self.mass[parent]
        omega = omega_deg * np.pi / 180
        mean_anomaly = mean_anomaly_deg * np.pi / 180
        if direction > 0:
            mean_anomaly = -mean_anomaly
        if ecc == 0:
            ecc = 0.00001
        if ecc >= 1:
            ecc = 0.95
        if ap_d:
            a = (pe_d + ap_d) / 2
            ecc = (ap_d / a) - 1
        else:
            a = - pe_d / (ecc - 1)
        b = a * math.sqrt(1 - ecc**2)
        f = math.sqrt(a**2 - b**2)
        if true_anomaly_deg:
            ta = true_anomaly_deg * np.pi / 180
            ea = math.acos((ecc + math.cos(ta)) / (1 + (ecc * math.cos(ta))))
        else:
            ea = solve_kepler_ell(ecc, mean_anomaly, 1e-10)
        pr_x = a * math.cos(ea) - f
        pr_y = b * math.sin(ea)
        pr = np.array([pr_x * math.cos(omega - np.pi) - pr_y * math.sin(omega - np.pi),
                       pr_x * math.sin(omega - np.pi) + pr_y * math.cos(omega - np.pi)])
        # velocity direction: implicit derivative of the rotated ellipse, disambiguated against
        # the chord through the curve's x-extremes (same construction as
        # convert.kepler_to_velocity); the speed uses the expression of :506
        so, co = math.sin(omega), math.cos(omega)
        p_x, p_y = pr[0] - f * np.cos(omega), pr[1] - f * np.sin(omega)
        vr_angle = np.arctan(
            -(b**2 * p_x * co**2 + a**2 * p_x * so**2 + b**2 * p_y * so * co - a**2 * p_y * so * co)
            / (a**2 * p_y * co**2 + b**2 * p_y * so**2 + b**2 * p_x * so * co - a**2 * p_x * so * co))
        x_max = math.sqrt(a**2 * math.cos(2*omega) + a**2 - b**2 * math.cos(2*omega) + b**2) / math.sqrt(2) - 10**-6
        x2, c2, s2 = x_max**2, co**2, so**2
        y_max = (a*b*math.sqrt(-x2 * (co**4 + so**4) - 2 * x2 * c2 * s2 + b**2 * s2 + a**2 * c2)
                 + a**2 * x_max * so * co - b**2 * x_max * so * co) / (a**2 * c2 + b**2 * s2)
        x1, y1, xx2, yy2 = x_max + f * co, y_max + f * so, -x_max + f * co, -y_max + f * so
        if ((xx2 - x1) * (pr[1] - y1)) - ((yy2 - y1) * (pr[0] - x1)) < 0:
            vr_angle += np.pi
        vr_angle = vr_angle % (2*np.pi)
        prm = math.sqrt(pr[0]*pr[0] + pr[1]*pr[1])
        vrm = direction * math.sqrt((2 * a * u - prm * u) / (a * prm))
        self.move_parent(body, self.pos[parent] + pr)
        self.rel_vel[body] = np.array([vrm * math.cos(vr_angle), vrm * math.sin(vr_angle)])
        for body_s in np.argsort(self.mass)[-1::-1]:
            self.vel[body_s] = self.rel_vel[body_s] + self.vel[self.parents[body_s]]

    # editor.py:1503-1523
    def frame(self, warp=1):
        deleted = []
        for _ in range(warp):
            self.gravity()
            self.body()
            d = self.inelastic_collision()
            if d is not None:
                deleted.append(int(d))
        self.kepler_basic()
        return deleted


class AllPairsOracle:
    """MODE_ALLPAIRS: documented complex model (01-complex-orbit-model.txt:4-12)
    + B1 detection + the documented merge (09-collisions.txt:3-6): survivor gets
    summed mass and momentum-weighted velocity, keeps its own position and
    density; the deleted row is removed np.delete-style (order preserved);
    rad is refreshed by body() at the start of the next tick's collision phase
    exactly like the editor order gravity(); body(); inelastic_collision()."""

    def __init__(self, mass, den, pos, vel, gc=1.0, rad_mult=179.0):
        self.gc, self.rad_mult = float(gc), float(rad_mult)
        self.mass, self.den = _f(mass).copy(), _f(den).copy()
        self.pos, self.vel = _f(pos).copy(), _f(vel).copy()
        self.rad = body_radius(self.mass, self.den, self.rad_mult)

    def gravity(self):
        self.pos, self.vel = allpairs_step(self.mass, self.pos, self.vel, self.gc, 1)

    def body(self):
        self.rad = body_radius(self.mass, self.den, self.rad_mult)

    def inelastic_collision(self):
        d, a = check_collision(self.mass, self.pos, self.rad)
        if d is None:
            return None, None
        mass_r = self.mass[d] + self.mass[a]
        velr = (self.vel[a] * self.mass[a] + self.vel[d] * self.mass[d]) / mass_r
        self.vel[a] = velr
        self.mass[a] = mass_r
        for name in ("mass", "den", "rad"):
            setattr(self, name, np.delete(getattr(self, name), d))
        for name in ("pos", "vel"):
            setattr(self, name, np.delete(getattr(self, name), d, axis=0))
        return d, a

    def tick(self):
        self.gravity()
        self.body()
        return self.inelastic_collision()
