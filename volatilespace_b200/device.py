"""Thin object layer over the C ABI: one `Device` per GPU context.

Nothing here computes physics; every method forwards numpy buffers to
libvsb200.so (see include/vsb200.h) and raises VsbError on failure.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import (F_ACC, F_COI, F_DEN, F_MASS, F_PARENTS, F_POS, F_RAD, F_REL_VEL, F_VEL,
                   KIND_BODY, KIND_VESSEL, check, f64, i64, ptr)

_default = None


def default_device():
    """Process-wide context on cuda:LOCAL_RANK (cuda:0 when not distributed)."""
    global _default
    if _default is None:
        import os
        _default = Device(int(os.environ.get("LOCAL_RANK", "0")))
    return _default


_scratch = None


def scratch_device():
    """A second context on the same GPU for the stateless free functions: several `vsb_host_*`
    calls stage their arrays in the context's body state, which must not disturb a live Physics
    object that shares the default context."""
    global _scratch
    if _scratch is None:
        _scratch = Device(default_device().index)
    return _scratch


def _unregister(L, address):
    try:
        L.vsb_host_unregister(C.c_void_p(address))
    except Exception:       # interpreter shutdown
        pass


def pinned_empty(shape, dtype=np.float64, L=None):
    """np.empty whose memory is page-locked with vsb_host_register (cudaHostRegister), so that the
    library's synchronous copies into / out of it run at PCIe rate instead of through the driver's
    pageable staging (~5 GB/s).  The registration is dropped when the buffer is garbage collected.
    Falls back to a plain array when the registration is refused (no device, locked-memory limit):
    pinning is an optimisation, never a requirement."""
    L = L or _lib.load()
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    raw = np.empty(max(nbytes, 1) + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    if nbytes >= 65536 and L.vsb_host_register(C.c_void_p(raw.ctypes.data), raw.nbytes) == _lib.VSB_OK:
        weakref.finalize(raw, _unregister, L, raw.ctypes.data)
    return raw[off:off + nbytes].view(dtype).reshape(shape)


class Device:
    def __init__(self, device=0):
        self.L = _lib.load()
        self.index = int(device)
        h = C.c_void_p()
        check(self.L.vsb_ctx_create(self.index, C.byref(h)))
        self.h = h
        self._pool = {}

    def staging(self, name, shape, dtype=np.float64):
        """Grow-only page-locked result buffer `name` of this Device, viewed as `shape`.  The view
        is valid until the next call that asks for the same name (the `reuse=True` results below:
        the callers of the reference consume them within the frame, game.py:1248-1296)."""
        shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
        count = int(np.prod(shape, dtype=np.int64))
        buf = self._pool.get(name)
        if buf is None or buf.dtype != np.dtype(dtype) or buf.size < count:
            buf = pinned_empty(count + count // 4 + 16, dtype, self.L)
            self._pool[name] = buf
        return buf[:count].reshape(shape)

    def close(self):
        if getattr(self, "h", None):
            self.L.vsb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------ misc
    def sync(self):
        check(self.L.vsb_ctx_sync(self.h))

    def stream_handle(self):
        s = C.c_void_p()
        check(self.L.vsb_ctx_stream(self.h, C.byref(s)))
        return s.value or 0

    def kernel_launches(self):
        n = C.c_int64()
        check(self.L.vsb_ctx_kernel_launches(self.h, C.byref(n)))
        return n.value

    def fp64_peak(self):
        """(FLOP/s, DFMA lane-ops/s) measured by a DFMA-chain microbenchmark."""
        a, b = C.c_double(), C.c_double()
        check(self.L.vsb_measure_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ------------------------------------------------------------ body-sim state
    def bodies_resize(self, n):
        check(self.L.vsb_bodies_resize(self.h, int(n)))

    def bodies_count(self):
        n = C.c_int64()
        check(self.L.vsb_bodies_count(self.h, C.byref(n)))
        return n.value

    def set(self, field, arr, first=0):
        w = _lib.BODY_FIELD_WIDTH[field]
        a = i64(arr) if field == F_PARENTS else f64(arr)
        count = a.size // w
        check(self.L.vsb_bodies_set(self.h, field, ptr(a), int(first), count))

    def get(self, field, first=0, count=None, out=None):
        w = _lib.BODY_FIELD_WIDTH[field]
        if count is None:
            count = self.bodies_count() - first
        dt = np.int64 if field == F_PARENTS else np.float64
        if out is None:
            out = np.empty((count, 2) if w == 2 else (count,), dtype=dt)
        assert out.dtype == dt and out.flags.c_contiguous and out.size == count * w
        check(self.L.vsb_bodies_get(self.h, field, ptr(out), int(first), int(count)))
        return out

    def device_ptr(self, field):
        p = C.c_void_p()
        check(self.L.vsb_bodies_device_ptr(self.h, field, C.byref(p)))
        return p.value

    def set_rows(self, row0, row1):
        check(self.L.vsb_bodies_set_rows(self.h, int(row0), int(row1)))

    def upload_bodies(self, mass, pos, vel=None, rad=None, coi=None, rel_vel=None, den=None,
                      parents=None):
        n = len(mass)
        self.bodies_resize(n)
        self.set(F_MASS, mass)
        self.set(F_POS, pos)
        for field, arr in ((F_VEL, vel), (F_RAD, rad), (F_COI, coi), (F_REL_VEL, rel_vel),
                           (F_DEN, den), (F_PARENTS, parents)):
            if arr is not None:
                self.set(field, arr)

    def find_parents(self, order, want=True):
        order = i64(order)
        out = np.empty(len(order), dtype=np.int64) if want else None
        check(self.L.vsb_find_parents(self.h, ptr(order), ptr(out)))
        return out

    def gravity_ref(self, gc, order):
        order = None if order is None else i64(order)
        check(self.L.vsb_gravity_ref(self.h, float(gc), ptr(order)))

    def gravity_ref_begin_async(self, order, part, parts):
        """First half of the MODE_REF tick with this rank's share of the find_parents scan."""
        order = None if order is None else i64(order)
        check(self.L.vsb_gravity_ref_begin_async(self.h, ptr(order), int(part), int(parts)))

    def find_parents_best_ptr(self):
        p = C.c_void_p()
        check(self.L.vsb_find_parents_best_ptr(self.h, C.byref(p)))
        return p.value

    def gravity_ref_finish(self, gc):
        check(self.L.vsb_gravity_ref_finish(self.h, float(gc)))

    def allpairs_accel(self, gc):
        check(self.L.vsb_allpairs_accel(self.h, float(gc)))

    def allpairs_integrate(self):
        check(self.L.vsb_allpairs_integrate(self.h))

    def allpairs_step(self, gc, nsteps=1):
        check(self.L.vsb_allpairs_step(self.h, float(gc), int(nsteps)))

    def set_gravity_mode(self, mode):
        """-1 / "auto", 0 / "rows": row kernel, 1 / "sym": symmetric kernel (include/vsb200.h)."""
        mode = {"auto": -1, "rows": 0, "sym": 1}.get(mode, mode)
        check(self.L.vsb_set_gravity_mode(self.h, int(mode)))

    def allpairs_sym_partial_async(self, gc, rank, world):
        check(self.L.vsb_allpairs_sym_partial_async(self.h, float(gc), int(rank), int(world)))

    def allpairs_integrate_async(self):
        check(self.L.vsb_allpairs_integrate_async(self.h))

    def allpairs_tick_async(self, gc):
        check(self.L.vsb_allpairs_tick_async(self.h, float(gc)))

    def editor_ticks(self, gc, coi_coef, order, max_ticks, do_kepler_basic=True, frame_out=None):
        """-> (ticks_done, body_del | None, body_add | None).  order=None reuses the order
        uploaded by the previous call; frame_out: float64 array of 4 + 14 n words that
        receives the frame record (see include/vsb200.h)."""
        order = None if order is None else i64(order)
        t, d, a = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.L.vsb_editor_ticks(self.h, float(gc), float(coi_coef), ptr(order),
                                      int(max_ticks), 1 if do_kepler_basic else 0, C.byref(t),
                                      C.byref(d), C.byref(a), ptr(frame_out)))
        if d.value < 0:
            return t.value, None, None
        return t.value, d.value, a.value

    def get_many(self, fields, out=None):
        """One D2H copy for several fields; returns the packed float64 buffer (parents'
        int64 words are bit-cast: use .view(np.int64) on that slice)."""
        n = self.bodies_count()
        words = sum(n * _lib.BODY_FIELD_WIDTH[f] for f in fields)
        if out is None:
            out = np.empty(words)
        arr = (C.c_int * len(fields))(*fields)
        check(self.L.vsb_bodies_get_many(self.h, arr, len(fields), ptr(out)))
        return out

    def kepler_basic(self, gc, coi_coef):
        check(self.L.vsb_kepler_basic(self.h, float(gc), float(coi_coef)))

    def simplified_orbit_coi(self, gc, coi_coef, order):
        order = i64(order)
        largest = C.c_int64()
        check(self.L.vsb_simplified_orbit_coi(self.h, float(gc), float(coi_coef), ptr(order),
                                              C.byref(largest)))
        return largest.value

    def check_collision(self):
        d, a = C.c_int64(), C.c_int64()
        check(self.L.vsb_check_collision(self.h, C.byref(d), C.byref(a)))
        if d.value < 0:
            return None, None
        return d.value, a.value

    def check_collision_part_async(self, part, parts):
        """Exhaustive scan of this rank's row blocks (multi-GPU); the local minimum key stays in
        the device word at collision_key_ptr()."""
        check(self.L.vsb_check_collision_part_async(self.h, int(part), int(parts)))

    def collision_key_ptr(self):
        p = C.c_void_p()
        check(self.L.vsb_collision_key_ptr(self.h, C.byref(p)))
        return p.value

    def check_collision_finish(self):
        d, a = C.c_int64(), C.c_int64()
        check(self.L.vsb_check_collision_finish(self.h, C.byref(d), C.byref(a)))
        if d.value < 0:
            return None, None
        return d.value, a.value

    def delete_row(self, row):
        check(self.L.vsb_delete_row(self.h, int(row)))

    def swap_rows(self, a, b):
        check(self.L.vsb_swap_rows(self.h, int(a), int(b)))

    def translate_to(self, row):
        check(self.L.vsb_translate_to(self.h, int(row)))

    def editor_curve(self, P, ell_t, par_t):
        n = self.bodies_count()
        out = np.empty((2, n, int(P)))
        ell_t, par_t = f64(ell_t), f64(par_t)
        check(self.L.vsb_editor_curve(self.h, int(P), ptr(ell_t), ptr(par_t), ptr(out)))
        return out

    # ------------------------------------------------------------ host-buffer calls
    def host_allpairs_step(self, mass, pos, vel, gc, nsteps=1):
        """In place on pos / vel (must be C-contiguous float64)."""
        assert pos.dtype == np.float64 and pos.flags.c_contiguous
        assert vel.dtype == np.float64 and vel.flags.c_contiguous
        mass = f64(mass)
        check(self.L.vsb_host_allpairs_step(self.h, len(mass), ptr(mass), ptr(pos), ptr(vel),
                                            float(gc), int(nsteps)))

    def host_allpairs_accel(self, mass, pos, gc):
        mass, pos = f64(mass), f64(pos)
        acc = np.empty((len(mass), 2))
        check(self.L.vsb_host_allpairs_accel(self.h, len(mass), ptr(mass), ptr(pos), float(gc),
                                             ptr(acc)))
        return acc

    def host_find_parents(self, order, pos, coi):
        order, pos, coi = i64(order), f64(pos), f64(coi)
        out = np.empty(len(order), dtype=np.int64)
        check(self.L.vsb_host_find_parents(self.h, len(order), ptr(order), ptr(pos), ptr(coi),
                                           ptr(out)))
        return out

    def host_check_collision(self, mass, pos, rad):
        mass, pos, rad = f64(mass), f64(pos), f64(rad)
        d, a = C.c_int64(), C.c_int64()
        check(self.L.vsb_host_check_collision(self.h, len(mass), ptr(mass), ptr(pos), ptr(rad),
                                              C.byref(d), C.byref(a)))
        if d.value < 0:
            return None, None
        return d.value, a.value

    def host_solve_kepler_ell(self, ecc, ma, tol=1e-10):
        ecc, ma = np.broadcast_arrays(f64(ecc), f64(ma))
        ecc, ma = f64(ecc), f64(ma)
        out = np.empty(ecc.shape)
        check(self.L.vsb_host_solve_kepler_ell(self.h, ecc.size, ptr(ecc), ptr(ma), float(tol),
                                               ptr(out)))
        return out

    def host_newton_root_kepler_hyp(self, ecc, ma, guess):
        ecc, ma, guess = (f64(x) for x in np.broadcast_arrays(f64(ecc), f64(ma), f64(guess)))
        out = np.empty(ecc.shape)
        check(self.L.vsb_host_newton_root_kepler_hyp(self.h, ecc.size, ptr(ecc), ptr(ma),
                                                     ptr(guess), ptr(out)))
        return out

    def host_curve_points(self, ecc, a, b, pea, t):
        ecc, a, b, pea, t = (f64(np.atleast_1d(x)) for x in (ecc, a, b, pea, t))
        out = np.empty((len(a), len(t), 2))
        check(self.L.vsb_host_curve_points(self.h, len(a), len(t), ptr(ecc), ptr(a), ptr(b),
                                           ptr(pea), ptr(t), ptr(out)))
        return out

    def host_curve_move_to(self, curves, body_pos, ref, f, pea):
        curves, body_pos, ref, f, pea = f64(curves), f64(body_pos), i64(ref), f64(f), f64(pea)
        out = np.empty_like(curves)
        check(self.L.vsb_host_curve_move_to(self.h, curves.shape[0], curves.shape[1],
                                            len(body_pos), ptr(curves), ptr(body_pos), ptr(ref),
                                            ptr(f), ptr(pea), ptr(out)))
        return out

    # ------------------------------------------------------------ Kepler state
    def kepler_load(self, kind, P, a, b, f, ecc, pea, n, dr, ma, ea, ref, t=None):
        arrs = [f64(x) for x in (a, b, f, ecc, pea, n, dr, ma, ea)]
        ref = i64(ref)
        t = np.linspace(-np.pi, np.pi, int(P)) if t is None else f64(t)
        check(self.L.vsb_kepler_load(self.h, kind, len(arrs[0]), int(P), *[ptr(x) for x in arrs],
                                     ptr(ref), ptr(t)))

    def kepler_set_order(self, order):
        order = i64(order)
        check(self.L.vsb_kepler_set_order(self.h, ptr(order)))

    def kepler_set(self, kind, field, arr, first=0):
        a = i64(arr) if field == _lib.K_REF else f64(arr)
        w = 2 if field == _lib.K_POS else 1
        check(self.L.vsb_kepler_set(self.h, kind, field, ptr(a), int(first), a.size // w))

    def kepler_get(self, kind, field, count, first=0, out=None):
        w = 2 if field == _lib.K_POS else 1
        dt = np.int64 if field == _lib.K_REF else np.float64
        if out is None:
            out = np.empty((count, 2) if w == 2 else (count,), dtype=dt)
        check(self.L.vsb_kepler_get(self.h, kind, field, ptr(out), int(first), int(count)))
        return out

    def kepler_move(self, kind, warp, body_pos=None):
        bp = None if body_pos is None else f64(body_pos)
        check(self.L.vsb_kepler_move(self.h, kind, float(warp), ptr(bp),
                                     0 if bp is None else len(bp)))

    def kepler_move_state(self, kind, warp, body_pos, pos, ma, ea):
        """move + (pos, ma, ea) of all objects into the given float64 arrays, one wait."""
        bp = None if body_pos is None else f64(body_pos)
        check(self.L.vsb_kepler_move_state(self.h, kind, float(warp), ptr(bp), 0 if bp is None else len(bp),
                                           ptr(pos), ptr(ma), ptr(ea)))

    def kepler_curve_move(self, kind, count, P, out=None):
        """curves of all objects + their copy to the host (count, P, 2), one wait."""
        if out is None:
            out = np.empty((count, P, 2))
        check(self.L.vsb_kepler_curve_move(self.h, kind, ptr(out)))
        return out

    def kepler_curves(self, kind):
        check(self.L.vsb_kepler_curves(self.h, kind))

    def kepler_tick(self, kind, warp, body_pos=None):
        bp = None if body_pos is None else f64(body_pos)
        check(self.L.vsb_kepler_tick(self.h, kind, float(warp), ptr(bp),
                                     0 if bp is None else len(bp)))

    def kepler_tick_async(self, kind, warp):
        check(self.L.vsb_kepler_tick_async(self.h, kind, float(warp)))

    def kepler_curves_get(self, kind, first, count, P, out=None):
        if out is None:
            out = np.empty((count, P, 2))
        check(self.L.vsb_kepler_curves_get(self.h, kind, int(first), int(count), ptr(out)))
        return out

    def host_body_thermal(self, consts12, mass, den, rad, atm_pres0, atm_scale_h, atm_den0, atm_h,
                          base_color):
        """N4; returns dict surf_grav, luminosity, temp, rad_sc, atm_h, color, types,
        stellar_class (codes 0..7)."""
        n = len(mass)
        ins = [f64(x) for x in (mass, den, rad, atm_pres0, atm_scale_h, atm_den0)]
        atm_h = f64(atm_h).copy()
        bc = i64(base_color)
        k = f64(consts12)
        out = {name: np.empty(n) for name in ("surf_grav", "luminosity", "temp", "rad_sc",
                                              "types")}
        color = np.empty((n, 3), dtype=np.int64)
        stellar = np.empty(n, dtype=np.int32)
        check(self.L.vsb_host_body_thermal(
            self.h, n, ptr(k), *[ptr(x) for x in ins], ptr(atm_h), ptr(bc), ptr(out["surf_grav"]),
            ptr(out["luminosity"]), ptr(out["temp"]), ptr(out["rad_sc"]), ptr(color),
            ptr(out["types"]), ptr(stellar)))
        out.update(atm_h=atm_h, color=color, stellar_class=stellar)
        return out

    def host_vessel_points(self, vessel12, v_ref, body12, body_ref, body_impact, body_enter_atm,
                           coi_leave, coi_enter, sel=None, entered_coi=None, left_coi=None,
                           left_coi_prev_ref=None, only_enter_coi=False, steps_limit=300):
        """phys_vessel.Physics.points (phys_vessel.py:372-474) for the vessels in `sel` (all when
        None); the four (nv,.) result arrays are updated in place."""
        vessel12, body12 = f64(vessel12), f64(body12)
        v_ref, body_ref = i64(v_ref), i64(body_ref)
        nv, nb = len(vessel12), len(body12)
        assert vessel12.shape == (nv, 12) and body12.shape == (nb, 12)
        for arr, w in ((body_impact, 2), (body_enter_atm, 2), (coi_leave, 2), (coi_enter, 6)):
            assert arr.dtype == np.float64 and arr.flags.c_contiguous and arr.shape == (nv, w)
        sel_arr = None if sel is None else i64(np.atleast_1d(sel))
        none = lambda x: -1 if x is None else int(x)
        check(self.L.vsb_host_vessel_points(
            self.h, nv, nb, ptr(vessel12), ptr(v_ref), ptr(body12), ptr(body_ref),
            0 if sel_arr is None else len(sel_arr), None if sel_arr is None else ptr(sel_arr),
            none(entered_coi), none(left_coi), none(left_coi_prev_ref), int(bool(only_enter_coi)),
            int(steps_limit), ptr(body_impact), ptr(body_enter_atm), ptr(coi_leave), ptr(coi_enter)))

    def host_check_warp(self, ecc, dr, ea, ma, n, body_impact, body_enter_atm, coi_leave, coi_enter,
                        warp, hold):
        """phys_vessel.Physics.check_warp (phys_vessel.py:506-560): `hold` (uint8 mask of
        physical_hold) is updated in place -> (situation, alarm_vessel, len(physical_hold))."""
        assert hold.dtype == np.uint8 and hold.flags.c_contiguous and len(hold) == len(ecc)
        out = np.zeros(3, dtype=np.int64)
        check(self.L.vsb_host_check_warp(
            self.h, len(ecc), ptr(f64(ecc)), ptr(f64(dr)), ptr(f64(ea)), ptr(f64(ma)), ptr(f64(n)),
            ptr(f64(body_impact)), ptr(f64(body_enter_atm)), ptr(f64(coi_leave)),
            ptr(f64(coi_enter)), float(warp), ptr(hold), ptr(out)))
        none = lambda x: None if x < 0 else int(x)
        return none(out[0]), none(out[1]), int(out[2])

    def host_enter_coi_service(self, dr, prev_ea, ea, coi_enter, crossing_pending=False):
        """predict_enter_coi_service (phys_vessel.py:744-758) -> indices it re-predicts."""
        flags = np.zeros(len(dr), dtype=np.uint8)
        check(self.L.vsb_host_enter_coi_service(
            self.h, len(dr), ptr(f64(dr)), ptr(f64(prev_ea)), ptr(f64(ea)), ptr(f64(coi_enter)),
            int(bool(crossing_pending)), ptr(flags)))
        return np.flatnonzero(flags)

    def host_cross_scan(self, ecc, dr, prev_ea, ea, body_impact, coi_leave, coi_enter):
        """cross_coi detection (phys_vessel.py:563-632) -> (vessel | None, kind 1 enter / 2 leave)."""
        out = np.zeros(2, dtype=np.int64)
        check(self.L.vsb_host_cross_scan(
            self.h, len(ecc), ptr(f64(ecc)), ptr(f64(dr)), ptr(f64(prev_ea)), ptr(f64(ea)),
            ptr(f64(body_impact)), ptr(f64(coi_leave)), ptr(f64(coi_enter)), ptr(out)))
        return (None if out[0] < 0 else int(out[0])), int(out[1])

    def host_cross_apply(self, kind, vessel8, ref, new_ref, body7, gc):
        """The orbit change of cross_coi for a batch of events -> (out6 (nev,6), status (nev,))."""
        vessel8, body7 = f64(np.atleast_2d(vessel8)), f64(body7)
        kind = np.ascontiguousarray(np.atleast_1d(kind), dtype=np.int32)
        ref, new_ref = i64(np.atleast_1d(ref)), i64(np.atleast_1d(new_ref))
        nev = len(kind)
        assert vessel8.shape == (nev, 8) and body7.shape[1] == 7
        out = np.empty((nev, 6))
        status = np.zeros(nev, dtype=np.int32)
        check(self.L.vsb_host_cross_apply(self.h, nev, ptr(kind), ptr(vessel8), ptr(ref),
                                          ptr(new_ref), len(body7), ptr(body7), float(gc), ptr(out),
                                          ptr(status)))
        return out, status

    # ---- resident vessel events (the KIND_VESSEL state must be loaded) ----
    def vessel_events_load(self, pe_d, ap_d, period, body_impact=None, body_enter_atm=None,
                           coi_leave=None, coi_enter=None, hold=None, prev_ea=None):
        opt = lambda a: None if a is None else ptr(f64(a))
        h = None if hold is None else np.ascontiguousarray(hold, dtype=np.uint8)
        check(self.L.vsb_vessel_events_load(self.h, opt(body_impact), opt(body_enter_atm),
                                            opt(coi_leave), opt(coi_enter), ptr(h), opt(prev_ea),
                                            ptr(f64(pe_d)), ptr(f64(ap_d)), ptr(f64(period))))

    def vessel_events_set_orbit(self, vessel, pe_d, ap_d, period):
        check(self.L.vsb_vessel_events_set_orbit(self.h, int(vessel), float(pe_d), float(ap_d),
                                                 float(period)))

    def vessel_events_get(self, n, first=0):
        """-> dict of the resident arrays for vessels [first, first+n)."""
        out = {"body_impact": np.empty((n, 2)), "body_enter_atm": np.empty((n, 2)),
               "coi_leave": np.empty((n, 2)), "coi_enter": np.empty((n, 6)),
               "hold": np.empty(n, dtype=np.uint8), "prev_ea": np.empty(n)}
        check(self.L.vsb_vessel_events_get(self.h, int(first), int(n), *(ptr(out[k]) for k in (
            "body_impact", "body_enter_atm", "coi_leave", "coi_enter", "hold", "prev_ea"))))
        return out

    def vessel_points(self, body12, body_ref, sel=None, entered_coi=None, left_coi=None,
                      left_coi_prev_ref=None, only_enter_coi=False, steps_limit=300):
        body12, body_ref = f64(body12), i64(body_ref)
        sel_arr = None if sel is None else i64(np.atleast_1d(sel))
        none = lambda x: -1 if x is None else int(x)
        check(self.L.vsb_vessel_points(
            self.h, len(body12), ptr(body12), ptr(body_ref),
            0 if sel_arr is None else len(sel_arr), None if sel_arr is None else ptr(sel_arr),
            none(entered_coi), none(left_coi), none(left_coi_prev_ref), int(bool(only_enter_coi)),
            int(steps_limit)))

    def vessel_check_warp(self, warp):
        out = np.zeros(3, dtype=np.int64)
        check(self.L.vsb_vessel_check_warp(self.h, float(warp), ptr(out)))
        none = lambda x: None if x < 0 else int(x)
        return none(out[0]), none(out[1]), int(out[2])

    def vessel_enter_coi_service(self, n, crossing_pending=False):
        idx = np.empty(n, dtype=np.int64)
        cnt = C.c_int64()
        check(self.L.vsb_vessel_enter_coi_service(self.h, int(bool(crossing_pending)), ptr(idx),
                                                  C.byref(cnt)))
        return idx[:cnt.value].copy()

    def vessel_cross_scan(self):
        out = np.zeros(2, dtype=np.int64)
        check(self.L.vsb_vessel_cross_scan(self.h, ptr(out)))
        return (None if out[0] < 0 else int(out[0])), int(out[1])

    def set_march_mode(self, sequential=False, exact_trig=True):
        """sequential: the reference's serial loop, one thread per pair; exact_trig: the march's
        sin / cos in double-double arithmetic, rounded to nearest (include/vsb200.h)."""
        check(self.L.vsb_set_march_mode(self.h, int(bool(sequential)) | (int(bool(exact_trig)) << 1)))

    def march_stats(self):
        """(chunks marched, chunks re-run, serial fallbacks, refused pairs) of the last COI-entry
        march."""
        out = np.zeros(4, dtype=np.uint64)
        check(self.L.vsb_march_stats(self.h, ptr(out)))
        return tuple(int(x) for x in out)

    def host_culling(self, coords, radius, screen_bounds, zoom):
        coords, radius, sb = f64(coords), f64(radius), f64(screen_bounds)
        mask = np.empty(len(radius), dtype=np.uint8)
        check(self.L.vsb_host_culling(self.h, len(radius), ptr(coords), ptr(radius), ptr(sb),
                                      float(zoom), ptr(mask)))
        return mask.astype(bool)

    def kepler_culling(self, kind, n, radius, screen_bounds, zoom):
        """`radius`: per-object array, None (zero), or one number for every object."""
        sb = f64(screen_bounds)
        idx = self.staging("cull_idx", n, np.int64)
        cnt = C.c_int64()
        if radius is not None and np.ndim(radius) == 0:
            check(self.L.vsb_kepler_culling_uniform(self.h, kind, float(radius), ptr(sb),
                                                    float(zoom), ptr(idx), C.byref(cnt)))
        else:
            radius = None if radius is None else f64(radius)
            check(self.L.vsb_kepler_culling(self.h, kind, ptr(radius), ptr(sb), float(zoom),
                                            ptr(idx), C.byref(cnt)))
        return idx[:cnt.value].copy()

    def kepler_visible_orbits(self, kind, n, ref_visible, ref_coi, size, zoom, screen_x):
        rv = np.ascontiguousarray(ref_visible, dtype=np.uint8)
        rc, size = f64(ref_coi), f64(size)
        idx = self.staging("cull_idx", n, np.int64)
        cnt = C.c_int64()
        check(self.L.vsb_kepler_visible_orbits(self.h, kind, len(rv), ptr(rv), ptr(rc), ptr(size),
                                               float(zoom), float(screen_x), ptr(idx),
                                               C.byref(cnt)))
        return idx[:cnt.value].copy()

    def kepler_curves_gather(self, kind, idx, P, reuse=False):
        """(len(idx), P, 2) moved curves of the listed objects.  reuse=True: the result is a view
        of this Device's page-locked staging buffer (valid until the next gather), which lets the
        copy run at PCIe rate and skips a fresh 16 P len(idx)-byte allocation per frame."""
        idx = i64(idx)
        shape = (len(idx), int(P), 2)
        out = self.staging("gather", shape) if reuse else np.empty(shape)
        check(self.L.vsb_kepler_curves_gather(self.h, kind, ptr(idx), len(idx), ptr(out)))
        return out

    def vessel_curve_segments(self, idx, P, reuse=False):
        """curve_segments for the listed vessels on the resident state -> (light, dark,
        first_intersect, intersect_type, intersect_time, select_range), one row per index.
        reuse=True: results are views of page-locked staging buffers (valid until the next call)."""
        idx = i64(idx)
        m = len(idx)
        new = (lambda name, shape, dt=np.float64: self.staging("seg_" + name, shape, dt)) if reuse \
            else (lambda name, shape, dt=np.float64: np.empty(shape, dtype=dt))
        light, dark = new("light", (m, int(P), 2)), new("dark", (m, int(P), 2))
        first, rng = new("first", (m, 2)), new("range", (m, 2))
        itype, itime = new("itype", m, np.int32), new("itime", m)
        check(self.L.vsb_vessel_curve_segments(self.h, ptr(idx), m, ptr(light), ptr(dark), ptr(first),
                                               ptr(itype), ptr(itime), ptr(rng)))
        return light, dark, first, itype, itime, rng

    def kepler_free(self, kind):
        check(self.L.vsb_kepler_free(self.h, kind))


__all__ = ["Device", "default_device", "KIND_BODY", "KIND_VESSEL", "F_MASS", "F_POS", "F_VEL",
           "F_RAD", "F_COI", "F_REL_VEL", "F_ACC", "F_DEN", "F_PARENTS"]
