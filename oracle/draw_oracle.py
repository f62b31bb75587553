"""CPU restatement of the drawing-side vessel functions of game mode (TEST INFRASTRUCTURE --
only tests/ may import this; the product path never does).

Follows the reference function by function (file:line of mzivic7/VolatileSpace v0.5.2):
  concat_wrap            phys_vessel.py:35-79
  curve_segments         phys_vessel.py:828-1075
  selected               phys_vessel.py:777-825
  rotate                 phys_vessel.py:761-774
  predict_next_orbit     phys_vessel.py:635-741
  orbit_time_to          phys_shared.py:58-64
  orb2xy                 phys_shared.py:187-199
  next_point / sort_intersect_indices   orbit_intersect.py:55-79
Scalar math goes through Python's `math` / numpy scalars, i.e. the same libm the reference
uses, so everything here is expected to match the live reference bit for bit
(tests/test_oracle_golden.py::test_vessel_draw_*, fixture tests/golden/vessel_draw.npz).

Vessel rows use the column order of the fixtures:
  v  = [a, b, f, ecc, pea, n, dr, ma, ea, prev_ea, u, pe_d, ap_d, period]
  p  = [impact0, impact1, atm0, atm1, leave0, leave1, enter0 .. enter5]
"""
import math

import numpy as np

import oracle as orc

TWO_PI = 2 * np.pi
A, B, F, ECC, PEA, N, DR, MA, EA, PREV_EA, U, PE_D, AP_D, PERIOD = range(14)
IMPACT, ATM, LEAVE, ENTER = 0, 2, 4, 6


def orbit_time_to(src, dst, ecc, dr, n):
    if ecc >= 1:
        return max(-dr * (dst - src) / n, 0.0)
    if dr < 0:
        return ((src - dst) % TWO_PI) / n
    return ((dst - src) % TWO_PI) / n


def orb2xy(a, b, f, ecc, pea, ref_pos, ea):
    if not ea:
        return np.array([0.0, 0.0])
    if ecc < 1:
        xn, yn = a * np.cos(ea) - f, b * np.sin(ea)
    else:
        xn, yn = a * np.cosh(ea) - f, b * np.sinh(ea)
    c, s = np.cos(pea - np.pi), np.sin(pea - np.pi)
    return np.array([(xn * c - yn * s) + ref_pos[0], (xn * s + yn * c) + ref_pos[1]])


def _angles_from(ea_vessel, pts, direction):
    ang = np.where(np.isnan(pts), np.nan, np.abs(ea_vessel - pts))
    ang = np.where(ea_vessel > pts, TWO_PI - ang, ang)
    return TWO_PI - ang if direction < 0 else ang


def sort_intersect(ea_vessel, pts, direction):
    """Indices of pts by angular distance ahead of the vessel; -1 for NaN entries; None if all NaN."""
    if np.all(np.isnan(pts)):
        return None
    order = np.argsort(_angles_from(ea_vessel, pts, direction))
    return np.where(np.isnan(pts[order]), -1, order)


def next_point(ea_vessel, pts, direction):
    if np.any(np.isnan(pts)):
        return np.array([np.nan, np.nan])
    srt = pts[np.argsort(_angles_from(ea_vessel, pts, direction))]
    return np.array([srt[0], srt[-1]])


def concat_wrap(array, p_first, r0, r1, p_last):
    """[p_first, array[r0:r1], p_last] in a NaN-padded array of the same shape; r0 > r1 walks
    through the end of the array and the end points swap places."""
    rows = array.shape[0]
    out = np.full((rows, 2), np.nan)
    room = rows - 2
    if r0 <= r1:
        cnt = min(r1 - r0, room)
        out[0] = p_first
        out[1:1 + cnt] = array[r0:r0 + cnt]
        out[1 + cnt] = p_last
    else:
        head, tail = array[r0:], array[:r1]
        cnt = min(len(head) + len(tail), room)
        n_head = min(cnt, len(head))
        out[0] = p_last
        out[1:1 + n_head] = head[:n_head]
        out[1 + n_head:1 + cnt] = tail[:cnt - n_head]
        out[1 + cnt] = p_first
    if np.any(np.isnan(array)):
        good = out[~np.isnan(out[:, 0])]
        out[:] = np.nan
        out[:len(good)] = good
    return out


def _index_ell(ea, P):
    return round(ea * P / TWO_PI)


def _index_hyp(ea, P):
    return P - round((ea + np.pi) * P / TWO_PI)


def _hyp_light(curve, ea, ev, en, ccw, here, there, clamp):
    """Hyperbola: the lit part between the vessel and the next point (the curve array runs from
    ea = +pi down to -pi, so index order is reversed)."""
    if ccw:
        if ev == en:
            ev += 1
        hi = max(ev - 1, 0) if clamp else ev - 1
        if clamp or ea < 0:
            return concat_wrap(curve, there, en, hi, here)
        return concat_wrap(curve, here, en, hi, there)
    if ev == en:
        ev -= 1
    if clamp or not ea < 0:
        return concat_wrap(curve, here, ev + 1, en, there)
    return concat_wrap(curve, there, ev + 1, en, here)


def curve_segments(P, v, p, pos, ref, body_pos, curves_mov, visible):
    """Returns (light, dark, dark_valid, first_intersect, intersect_type, intersect_time,
    select_range); rows of vessels outside `visible` are NaN / 0.  dark_valid marks the rows of
    `dark` the reference writes (it leaves np.empty garbage when a vessel has no point at all)."""
    nv = len(v)
    light = np.full((nv, P, 2), np.nan)
    dark = np.full((nv, P, 2), np.nan)
    dark_valid = np.zeros(nv, dtype=bool)
    first = np.full((nv, 2), np.nan)
    itype = np.zeros(nv, dtype=np.int64)
    itime = np.zeros(nv)
    srange = np.full((nv, 2), np.nan)
    for i in visible:
        a, b, f, ecc, pea, n, dr, ma, ea = v[i, [A, B, F, ECC, PEA, N, DR, MA, EA]]
        ell, ccw = ecc < 1, dr > 0
        col = 0 if ell else 1
        here = pos[i]
        curve = np.array(curves_mov[i])
        rp = body_pos[ref[i]]
        coord = lambda e: orb2xy(a, b, f, ecc, pea, rp, e)
        index = (lambda e: _index_ell(e, P)) if ell else (lambda e: _index_hyp(e, P))
        cand = np.array([p[i, IMPACT + col], p[i, ENTER + 1], p[i, LEAVE + col]])
        order = sort_intersect(ea, cand, dr)
        if order is None:
            light[i] = curve
            continue
        nearest = order[0]
        dark_valid[i] = True
        if nearest in (0, 1):
            # impact (0) or COI entry (1): the far part is the whole curve unless a COI leave cuts it
            dark[i] = curve
            ea_next = p[i, IMPACT] if nearest == 0 else p[i, ENTER + 1]
            there = coord(ea_next)
            ev, en = index(ea), index(ea_next)
            if not ell:
                light[i] = _hyp_light(curve, ea, ev, en, ccw, here, there, False)
            elif nearest == 0:
                if ccw:
                    if ev == en:
                        ev -= 1
                    light[i] = concat_wrap(curve, here, ev + 1, en, there)
                else:
                    if ev == en:
                        ev += 1
                    light[i] = concat_wrap(curve, there, en, ev - 1, here)
            elif ccw:
                if ea < np.pi:
                    if ev == en:
                        ev -= 1
                    light[i] = concat_wrap(curve, here, ev + 1, en, there)
                elif np.pi < ea < np.pi * 1.5:
                    light[i] = concat_wrap(curve, here, ev + 1, en, there)
                else:
                    light[i] = concat_wrap(curve, there, ev, en, here)
            else:
                if not ea < np.pi and ev == en:
                    ev += 1
                light[i] = concat_wrap(curve, there, en, ev - 1, here)
            first[i] = there
            itime[i] = orbit_time_to(ma, ea_next - ecc * np.sin(ea_next), ecc, dr, n)
            itype[i] = 1 + nearest
            if nearest == 1:
                srange[i] = [ea, ea_next]
        if nearest == 2:
            # leaving the COI is the next thing that happens
            ea_next, ea_prev = p[i, LEAVE], p[i, LEAVE + 1]
            there, behind = coord(ea_next), coord(ea_prev)
            ev, en, ep = index(ea), index(ea_next), index(ea_prev)
            if ell:
                ma_next = ea_next - ecc * np.sin(ea_next)
                if ccw:
                    if ea < np.pi:
                        if ev == en:
                            ev -= 1
                        light[i] = concat_wrap(curve, here, ev + 1, en, there)
                    else:
                        light[i] = concat_wrap(curve, there, ev, en, here)
                    dark[i] = concat_wrap(curve, there, ep, en, behind)
                else:
                    if ea < np.pi:
                        light[i] = concat_wrap(curve, here, en, max(ev, 0), there)
                    else:
                        if ev == en:
                            ev += 1
                        light[i] = concat_wrap(curve, there, en, max(ev - 1, 0), here)
                    dark[i] = concat_wrap(curve, behind, en, ep, there)
            else:
                ma_next = (ecc * np.sinh(ea_next) - ea_next) * -1
                light[i] = _hyp_light(curve, ea, ev, en, ccw, here, there, True)
                if ccw:
                    dark[i] = concat_wrap(curve, there, en, ep, behind)
                else:
                    dark[i] = concat_wrap(curve, behind, ep, en, there)
            first[i] = there
            itime[i] = orbit_time_to(ma, ma_next, ecc, dr, n)
            itype[i] = 3
        elif np.any(order == 2):
            # something else comes first, the COI leave still ends the far part
            ea_next, ea_prev = p[i, LEAVE], p[i, LEAVE + 1]
            there, behind = coord(ea_next), coord(ea_prev)
            en, ep = index(ea_next), index(ea_prev)
            if ell:
                if ccw:
                    dark[i] = concat_wrap(curve, there, ep, en, behind)
                elif ea < np.pi:
                    dark[i] = concat_wrap(curve, behind, en, ep, there)
                else:
                    dark[i] = concat_wrap(curve, there, en, ep, behind)
            elif ccw:
                dark[i] = concat_wrap(curve, there, en, ep, behind)
            else:
                dark[i] = concat_wrap(curve, behind, ep, en, there)
    return light, dark, dark_valid, first, itype, itime, srange


def selected(vrow, pos_v, pos_ref):
    """[ta, pe_x, pe_y, pe_t, ap_x, ap_y, ap_t, distance, speed_orb, speed_hor, speed_vert]"""
    a, b, ecc, pea, n, dr, ma, ea, u, pe_d, ap_d = (float(vrow[k]) for k in
                                                   (A, B, ECC, PEA, N, DR, MA, EA, U, PE_D, AP_D))
    rel = pos_v - pos_ref
    dist = math.sqrt(rel[0] * rel[0] + rel[1] * rel[1])
    speed = math.sqrt(u * ((2 / dist) - (1 / a)))
    c = math.cos(ea) if ecc < 1 else math.cosh(ea)
    ta = math.acos((c - ecc) / (1 - ecc * c))
    if ma > np.pi:
        ta = TWO_PI - ta
    angle = (ta - math.atan(-b * (math.cos(ea)) / (math.sin(ea) * a))) % np.pi
    vert, hor = speed * math.cos(angle), speed * math.sin(angle)
    ap, ap_t = np.array([0.0, 0.0]), 0.0
    if ecc != 0 and ecc < 1:
        ap = np.array([ap_d * math.cos(pea), ap_d * math.sin(pea)]) + pos_ref
        ap_t = orbit_time_to(ma, np.pi, ecc, dr, n)
    if ecc == 0:
        ma = ta
    pe_t = orbit_time_to(ma, 0, ecc, dr, n)
    pe = np.array([pe_d * math.cos(pea - np.pi), pe_d * math.sin(pea - np.pi)]) + pos_ref
    return np.array([ta, pe[0], pe[1], pe_t, ap[0], ap[1], ap_t, dist, speed, hor, vert])


def rotate(warp, vessel, direction, rot_angle, rot_speed, rot_acc):
    """Returns (rot_angle, rot_speed) after the call."""
    rot_angle, rot_speed = np.array(rot_angle, dtype=float), np.array(rot_speed, dtype=float)
    if warp == 1:
        rot_speed[vessel] += rot_acc[vessel] * direction
        rot_angle = rot_angle + rot_speed
        rot_angle = np.where(rot_angle > TWO_PI, 0, rot_angle)
        rot_angle = np.where(rot_angle < 0, TWO_PI, rot_angle)
    else:
        rot_speed = rot_speed * 0.0
    return rot_angle, rot_speed


def _advance(ma, ecc, dr, n, dt):
    """phys_vessel.move :113-127 (its two wrap statements have no effect)."""
    if ecc < 1:
        ma = ma + dr * n * dt
        return ma, float(orc.solve_kepler_ell(np.array([ecc]), np.array([ma]))[0])
    ma = ma + dr * n * -1 * dt
    return ma, float(orc.newton_root_kepler_hyp(np.array([ecc]), np.array([ma]), np.array([ma]))[0])


def predict_next_orbit(i, P, v, p, ref, body_static, body_mass, body_ref, body_ma, gc, dt, t):
    """None, or (curve, [ap_x, ap_y, ap_d, ap_t, pe_x, pe_y, pe_d, pe_t, new_ref, -ref_pos(2),
    -vessel_pos(2), enter_coi]).  body_static columns: a, b, f, ecc, pea, n, dr, coi."""
    a, b, f, ecc, pea, dr, u = (float(v[i, k]) for k in (A, B, F, ECC, PEA, DR, U))
    zero = np.array((0.0, 0.0))
    bs = body_static

    def body_xy(k, origin, ea):
        return orb2xy(bs[k, 0], bs[k, 1], bs[k, 2], bs[k, 3], bs[k, 4], origin, ea)

    enter = p[i, ENTER:ENTER + 6]
    if not np.isnan(enter[1]):
        old, new = int(ref[i]), int(enter[0])
        if not dt or old == new:
            return None
        new_u = gc * body_mass[new]
        ref_then = body_xy(new, zero, enter[2])
        ves_then = orb2xy(a, b, f, ecc, pea, zero, enter[1])
        vel = orc.kepler_to_velocity(ves_then, a, ecc, pea, u, dr)
        ref_vel = orc.kepler_to_velocity(ref_then, bs[new, 0], bs[new, 3], bs[new, 4], new_u, bs[new, 6])
        el = orc.velocity_to_kepler(ves_then - ref_then, vel - ref_vel, new_u, failsafe=2)
        entering = True
    elif not np.isnan(p[i, LEAVE]) and np.isnan(p[i, IMPACT]):
        old = int(ref[i])
        new = int(body_ref[old])
        if not dt:
            return None
        _, ref_ea = _advance(body_ma[old], bs[old, 3], bs[old, 6], bs[old, 5], dt)
        ref_then = zero
        old_then = body_xy(old, ref_then, ref_ea)
        new_u = gc * body_mass[new]
        ves_then = orb2xy(a, b, f, ecc, pea, old_then, p[i, LEAVE])
        rel = ves_then - old_then
        vel = orc.kepler_to_velocity(rel, a, ecc, pea, u, dr)
        old_vel = orc.kepler_to_velocity(old_then, bs[old, 0], bs[old, 3], bs[old, 4], new_u, bs[old, 6])
        el = orc.velocity_to_kepler(rel + old_then, vel + old_vel, new_u, failsafe=1)
        entering = False
    else:
        return None
    na, necc, npea, nma, ndr = (float(x) for x in el)
    pea_inv = npea
    npea = npea + np.pi
    if npea >= TWO_PI:
        npea -= TWO_PI
    co = orc.calc_orb("vessel", [0], [body_mass[new]], [0.0], gc, 0.0, [na], [necc])
    nb, nf, pe_d, ap_d, nn = (float(co[k][0]) for k in ("b", "f", "pe_d", "ap_d", "n"))
    curve = orc.curve_points(necc, na, nb, npea, t)
    curve = curve + np.column_stack((nf * np.cos(npea), nf * np.sin(npea))) - ref_then
    ell = necc < 1
    coi = bs[new, 7]
    if not ell or ap_d > coi:
        cut = orc.ell_hyp_intersect_circle(na, nb, necc, nf, 0.0, coi)
        if ell:
            ea1 = float(orc.solve_kepler_ell(np.array([necc]), np.array([nma + nn * ndr]))[0])
        else:
            m1 = nma + nn * ndr * -1
            ea1 = float(orc.newton_root_kepler_hyp(np.array([necc]), np.array([m1]), np.array([m1]))[0])
        ea_next, ea_prev = next_point(ea1, cut, ndr)
        if not np.isnan(ea_next):
            there = -orb2xy(na, nb, nf, necc, pea_inv, ref_then, ea_next)
            behind = -orb2xy(na, nb, nf, necc, pea_inv, ref_then, ea_prev)
            if ell:
                en, ep = _index_ell(ea_next, P), _index_ell(ea_prev, P)
                curve = concat_wrap(curve, there, ep, en, behind) if ndr > 0 else \
                    concat_wrap(curve, behind, en, ep, there)
            else:
                en, ep = _index_hyp(ea_next, P), _index_hyp(ea_prev, P)
                curve = concat_wrap(curve, there, en, ep, behind) if ndr > 0 else \
                    concat_wrap(curve, behind, ep, en, there)
    pe_t = orbit_time_to(nma, 0, necc, ndr, nn) + dt
    if necc != 0 and necc < 1:
        ap = np.array([ap_d * math.cos(npea), ap_d * math.sin(npea)]) - ref_then
        ap_t = orbit_time_to(nma, np.pi, necc, ndr, nn) + dt
    else:
        ap, ap_t = np.array([0.0, 0.0]), 0.0
    pe = np.array([pe_d * math.cos(npea - np.pi), pe_d * math.sin(npea - np.pi)]) - ref_then
    sc = np.array([ap[0], ap[1], ap_d, ap_t, pe[0], pe[1], pe_d, pe_t, float(new), -ref_then[0],
                   -ref_then[1], -ves_then[0], -ves_then[1], float(entering)])
    return curve, sc
