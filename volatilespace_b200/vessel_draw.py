"""Host side of the drawing helpers of volatilespace/physics/phys_vessel.py that concern ONE
vessel per frame (the active / selected one): `selected` (:777-825), `rotate` (:761-774) and
`predict_next_orbit` (:635-741).  They are O(1) or O(P) per frame by construction, so they stay
on the host like the reference's own scalar code; the conic arithmetic inside
predict_next_orbit goes through the same batched library calls the other shims use
(orb2xy, kepler_to_velocity, velocity_to_kepler, ell_hyp_intersect_circle, the Kepler solvers,
curve_points).  `curve_segments`, the data-parallel one, is a kernel (csrc/vsb_segments.cu).
"""
import math

import numpy as np

from . import convert, enhanced_kepler_solver, orbit_intersect, phys_shared
from .synth import calc_orb

TAU = 2 * np.pi


def orbit_time_to(source_angle, target_angle, ecc, dr, n):
    """phys_shared.py:58-64."""
    if ecc >= 1:
        return max(-dr * (target_angle - source_angle) / n, 0.0)
    span = (source_angle - target_angle) if dr < 0 else (target_angle - source_angle)
    return (span % TAU) / n


def next_point(ea_vessel, ea_points, direction):
    """orbit_intersect.py:55-65: (nearest, farthest) of ea_points ahead of the vessel."""
    ea_points = np.asarray(ea_points, dtype=np.float64)
    if np.isnan(ea_points).any():
        return np.array([np.nan, np.nan])
    gap = np.abs(ea_vessel - ea_points)
    gap = np.where(ea_vessel > ea_points, TAU - gap, gap)
    if direction < 0:
        gap = TAU - gap
    ranked = ea_points[np.argsort(gap)]
    return np.array([ranked[0], ranked[-1]])


def concat_wrap(array, point_start, range_start, range_end, point_end):
    """phys_vessel.py:35-79 for one (P,2) curve (the batched form is vsb_vessel_curve_segments)."""
    rows = len(array)
    out = np.full((rows, 2), np.nan)
    if range_start <= range_end:
        body = array[range_start:range_start + min(range_end - range_start, rows - 2)]
        ends = (point_start, point_end)
    else:
        body = np.concatenate([array[range_start:], array[:range_end]])[:rows - 2]
        ends = (point_end, point_start)
    out[0] = ends[0]
    out[1:1 + len(body)] = body
    out[1 + len(body)] = ends[1]
    if np.isnan(array).any():
        kept = out[~np.isnan(out[:, 0])]
        out[:] = np.nan
        out[:len(kept)] = kept
    return out


def selected(el, pos_vessel, pos_ref):
    """phys_vessel.py:777-825.  el: dict with a, b, ecc, pea, n, dr, ma, ea, u, pe_d, ap_d of the
    vessel.  Returns (ta, pe, pe_t, ap, ap_t, distance, speed_orb, speed_hor, speed_vert)."""
    a, b, ecc, pea, ea, ma = el["a"], el["b"], el["ecc"], el["pea"], el["ea"], el["ma"]
    rel = np.asarray(pos_vessel, dtype=np.float64) - pos_ref
    distance = math.sqrt(rel[0] * rel[0] + rel[1] * rel[1])
    speed_orb = math.sqrt(el["u"] * ((2 / distance) - (1 / a)))     # vis-viva
    ch = math.cos(ea) if ecc < 1 else math.cosh(ea)
    ta = math.acos((ch - ecc) / (1 - ecc * ch))
    if ma > np.pi:
        ta = TAU - ta
    flight = (ta - math.atan(-b * (math.cos(ea)) / (math.sin(ea) * a))) % np.pi
    speed_vert, speed_hor = speed_orb * math.cos(flight), speed_orb * math.sin(flight)
    ap, ap_t = np.array([0, 0]), 0
    if 0 < ecc < 1 or ecc < 0:
        ap = np.array([el["ap_d"] * math.cos(pea), el["ap_d"] * math.sin(pea)]) + pos_ref
        ap_t = orbit_time_to(ma, np.pi, ecc, el["dr"], el["n"])
    pe_t = orbit_time_to(ta if ecc == 0 else ma, 0, ecc, el["dr"], el["n"])
    pe = np.array([el["pe_d"] * math.cos(pea - np.pi), el["pe_d"] * math.sin(pea - np.pi)]) + pos_ref
    return ta, pe, pe_t, ap, ap_t, distance, speed_orb, speed_hor, speed_vert


def rotate(warp, vessel, direction, rot_angle, rot_speed, rot_acc):
    """phys_vessel.py:761-774 -> new rot_angle (rot_speed is updated in place)."""
    if warp != 1:
        rot_speed *= 0.0
        return rot_angle
    rot_speed[vessel] += rot_acc[vessel] * direction
    rot_angle = rot_angle + rot_speed
    rot_angle = np.where(rot_angle > TAU, 0, rot_angle)
    return np.where(rot_angle < 0, TAU, rot_angle)


_NOTHING = (None,) * 11


def predict_next_orbit(ph, vessel):
    """phys_vessel.py:635-741 for the vessel shim `ph` (needs points() and curve_segments())."""
    enter = ph.coi_enter[vessel]
    a, b, f, ecc, pea, dr, u = (float(getattr(ph, k)[vessel]) for k in
                                ("a", "b", "f", "ecc", "pea", "dr", "u"))
    origin = np.array((0.0, 0.0))

    def body_at(k, centre, ea):
        return phys_shared.orb2xy(ph.body_a[k], ph.body_b[k], ph.body_f[k], ph.body_ecc[k],
                                  ph.body_pea[k], centre, ea)

    def body_velocity(k, at, mu):
        return convert.kepler_to_velocity(at, ph.body_a[k], ph.body_ecc[k], ph.body_pea[k], mu,
                                          ph.body_dr[k])

    dt = ph.intersect_time[vessel]
    if not np.isnan(enter[1]):
        old, new = int(ph.ref[vessel]), int(enter[0])
        if not dt or old == new:
            return _NOTHING
        mu = ph.gc * ph.body_mass[new]
        centre = body_at(new, origin, enter[2])                    # the new parent, then
        ves_abs = phys_shared.orb2xy(a, b, f, ecc, pea, origin, enter[1])
        vel = convert.kepler_to_velocity(ves_abs, a, ecc, pea, u, dr) - body_velocity(new, centre, mu)
        new_el = convert.velocity_to_kepler(ves_abs - centre, vel, mu, failsafe=2)
        entering = True
    elif not np.isnan(ph.coi_leave[vessel, 0]) and np.isnan(ph.body_impact[vessel, 0]):
        old = int(ph.ref[vessel])
        new = int(ph.body_ref[old])
        if not dt:
            return _NOTHING
        # where the body being left will be by then (phys_vessel.move :113-127)
        if ph.body_ecc[old] < 1:
            m_then = ph.body_ma[old] + ph.body_dr[old] * ph.body_n[old] * dt
            ea_then = enhanced_kepler_solver.solve_kepler_ell(float(ph.body_ecc[old]), float(m_then), 1e-10)
        else:
            m_then = ph.body_ma[old] + ph.body_dr[old] * ph.body_n[old] * -1 * dt
            ea_then = phys_shared.newton_root_kepler_hyp(float(ph.body_ecc[old]), float(m_then), float(m_then))
        centre = origin
        old_at = body_at(old, centre, ea_then)
        mu = ph.gc * ph.body_mass[new]
        ves_abs = phys_shared.orb2xy(a, b, f, ecc, pea, old_at, ph.coi_leave[vessel, 0])
        rel = ves_abs - old_at
        vel = convert.kepler_to_velocity(rel, a, ecc, pea, u, dr) + body_velocity(old, old_at, mu)
        new_el = convert.velocity_to_kepler(rel + old_at, vel, mu, failsafe=1)
        entering = False
    else:
        return _NOTHING
    a2, ecc2, pea_inv, ma2, dr2 = (float(x) for x in new_el)
    pea2 = pea_inv + np.pi
    if pea2 >= TAU:
        pea2 -= TAU
    orb = calc_orb("vessel", [0], [ph.body_mass[new]], None, ph.gc, 0.0, [a2], [ecc2])
    b2, f2, pe_d, ap_d, n2 = (float(orb[k][0]) for k in ("b", "f", "pe_d", "ap_d", "n"))
    curve = phys_shared.curve_points(ecc2, a2, b2, pea2, ph.t)
    curve = curve + np.column_stack((f2 * np.cos(pea2), f2 * np.sin(pea2))) - centre
    P = ph.curve_points
    closed = ecc2 < 1
    coi = ph.body_coi[new]
    if not closed or ap_d > coi:
        # cut the curve where it leaves the new parent's COI
        cut = orbit_intersect.ell_hyp_intersect_circle(a2, b2, ecc2, f2, 0.0, float(coi))
        if closed:
            ea_soon = enhanced_kepler_solver.solve_kepler_ell(ecc2, ma2 + n2 * dr2, 1e-10)
        else:
            m_soon = ma2 + n2 * dr2 * -1
            ea_soon = phys_shared.newton_root_kepler_hyp(ecc2, m_soon, m_soon)
        ea_next, ea_prev = next_point(ea_soon, cut, dr2)
        if not np.isnan(ea_next):
            ahead = -phys_shared.orb2xy(a2, b2, f2, ecc2, pea_inv, centre, ea_next)
            behind = -phys_shared.orb2xy(a2, b2, f2, ecc2, pea_inv, centre, ea_prev)
            if closed:
                k_next, k_prev = round(ea_next * P / TAU), round(ea_prev * P / TAU)
                curve = concat_wrap(curve, ahead, k_prev, k_next, behind) if dr2 > 0 else \
                    concat_wrap(curve, behind, k_next, k_prev, ahead)
            else:
                k_next = P - round((ea_next + np.pi) * P / TAU)
                k_prev = P - round((ea_prev + np.pi) * P / TAU)
                curve = concat_wrap(curve, ahead, k_next, k_prev, behind) if dr2 > 0 else \
                    concat_wrap(curve, behind, k_prev, k_next, ahead)
    pe_t = orbit_time_to(ma2, 0, ecc2, dr2, n2) + dt
    ap, ap_t = np.array([0, 0]), 0
    if ecc2 != 0 and closed:
        ap = np.array([ap_d * math.cos(pea2), ap_d * math.sin(pea2)]) - centre
        ap_t = orbit_time_to(ma2, np.pi, ecc2, dr2, n2) + dt
    pe = np.array([pe_d * math.cos(pea2 - np.pi), pe_d * math.sin(pea2 - np.pi)]) - centre
    return curve, ap, ap_d, ap_t, pe, pe_d, pe_t, new, -centre, -ves_abs, entering
