"""Generate tests/golden/*.npz from the LIVE reference (dev container only).

Run:  python oracle/gen_golden.py
Needs /root/reference (numba path, fastmath=False).  The reference ships no
golden vectors of its own, so these files -- outputs of the
reference itself on seeded inputs -- are what pins oracle/vs_oracle.c and,
through it, the CUDA path.  Everything is seeded; re-running reproduces the
files bit for bit on the same numpy/numba/libm.
"""
import os
import sys
from itertools import repeat

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def cloud(rng, n, extent, central_mass=1e6, vscale=1.0):
    """Small random 2-D system: heavy body 0 at origin + light bodies on roughly
    circular orbits (distinct masses)."""
    mass = rng.uniform(0.5, 2.0, n)
    mass[0] = central_mass
    r = rng.uniform(0.05, 1.0, n) * extent
    th = rng.uniform(0, 2 * np.pi, n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    sp = np.sqrt(central_mass / r) * rng.normal(1.0, 0.1, n) * vscale
    vel = np.stack([-sp * np.sin(th), sp * np.cos(th)], axis=1)
    pos[0] = 0
    vel[0] = 0
    den = np.full(n, 3000.0)
    return mass, den, pos, vel


def snap(ph, prefix):
    return {prefix + k: np.array(getattr(ph, k)) for k in
            ("mass", "rad", "pos", "vel", "rel_vel", "coi", "parents", "focus", "ecc_v",
             "semi_major", "semi_minor", "periapsis_arg")}


def gen_solar_inputs(ns):
    gd, conf, bd, bod, vd, vod = ns.peripherals.load_file(ns.solar_system_ini)
    return conf, bd, bod


def gen_solar(ns):
    gd, conf, bd, bod, vd, vod = ns.peripherals.load_file(ns.solar_system_ini)
    out = dict(mass=bd["mass"].copy(), den=bd["den"].copy(), pos0=bod["pos"].copy(),
               vel0=bod["vel"].copy(), gc=conf["gc"], rad_mult=conf["rad_mult"],
               coi_coef=conf["coi_coef"])
    ph = ns.phys_editor.Physics()
    ph.load_system(conf, bd, bod)
    out.update(load_coi=ph.coi.copy(), load_parents=np.array(ph.parents),
               load_rel_vel=ph.rel_vel.copy(), load_rad=ph.rad.copy(),
               load_largest=int(ph.largest))
    ph.kepler_basic(); ph.temp_color(); ph.classify()
    out.update(snap(ph, "f0_"))
    rh.set_curve_points(ns, ph, 64)
    frames = 0
    for target in (1, 2, 100, 1000, 10000):
        while frames < target:
            rh.editor_frame(ph)
            frames += 1
        out.update(snap(ph, "f%d_" % target))
        if target == 100:
            out["f100_curve64"] = ph.curve()
    save("solar_editor", **out)
    return conf, bd, bod


def gen_editor_clouds(ns):
    """MODE_REF frames with parent changes and merges on small seeded clouds."""
    rng = np.random.default_rng(20240611)
    out = {}
    # case A: 48-body hierarchical system (moons inside planet COIs), no overlap
    n = 48
    mass, den, pos, vel = cloud(rng, n, 2.0e4, central_mass=1e4)
    mass[1:7] = rng.uniform(50, 400, 6)                 # planets
    for k in range(7, 19):                               # moons near planets
        p = 1 + (k % 6)
        d = rng.uniform(40, 120)
        th = rng.uniform(0, 2 * np.pi)
        pos[k] = pos[p] + d * np.array([np.cos(th), np.sin(th)])
        sp = np.sqrt(mass[p] / d)
        vel[k] = vel[p] + sp * np.array([-np.sin(th), np.cos(th)])
        mass[k] = rng.uniform(0.01, 0.5)
    for k in range(19, 33):                              # COI crossers: fly-bys and escapers
        p = 1 + (k % 6)
        coi_est = np.hypot(*pos[p]) * (mass[p] / mass[0]) ** 0.7
        th = rng.uniform(0, 2 * np.pi)
        d = coi_est * (rng.uniform(1.02, 1.2) if k % 2 else rng.uniform(0.5, 0.9))
        u = np.array([np.cos(th), np.sin(th)])
        pos[k] = pos[p] + d * u
        sp = np.sqrt(mass[p] / d) * (rng.uniform(0.3, 1.2) if k % 2 else rng.uniform(1.5, 2.5))
        vel[k] = vel[p] - sp * u * (1 if k % 2 else -1) + 0.3 * sp * np.array([-u[1], u[0]])
        mass[k] = rng.uniform(0.01, 0.5)
    den[:] = 3.0e7                                       # tiny radii: case A is about COI crossings
    ph = rh.editor_from_arrays(ns, mass, den, pos, vel)
    out.update(A_mass=mass, A_den=den, A_pos0=pos, A_vel0=vel)
    out.update(snap(ph, "A_f0_"))
    dels, plog = [], []
    for fr in range(1, 601):
        d = rh.editor_frame(ph)
        dels.append(d[0] if d else -1)
        plog.append(np.array(ph.parents) if len(ph.parents) == n else np.full(n, -1))
        if fr in (1, 50, 300, 600):
            out.update(snap(ph, "A_f%d_" % fr))
    out["A_deleted"] = np.array(dels)
    out["A_parents_log"] = np.array([p if len(p) == n else np.full(n, -1) for p in plog])

    # case B: dense 40-body clump => merges (incl. index shifts)
    n = 40
    mass, den, pos, vel = cloud(rng, n, 600.0, central_mass=5e3, vscale=0.6)
    den[:] = 300.0
    ph = rh.editor_from_arrays(ns, mass, den, pos, vel)
    out.update(B_mass=mass, B_den=den, B_pos0=pos, B_vel0=vel)
    dels, counts = [], []
    for fr in range(1, 121):
        d = rh.editor_frame(ph)
        dels.append(d[0] if d else -1)
        counts.append(len(ph.mass))
        if fr in (1, 10, 40, 120):
            out.update(snap(ph, "B_f%d_" % fr))
    out["B_deleted"] = np.array(dels)
    out["B_counts"] = np.array(counts)

    # case C: the 5-body toy of SURVEY 8(c): two simultaneous overlaps (1,4),(2,3)
    mass = np.array([1e4, 5.0, 7.0, 3.0, 9.0])
    den = np.full(5, 1.0)
    pos = np.array([[0, 0], [5000.0, 0], [0, 8000.0], [0, 8010.0], [5010.0, 0]])
    vel = np.array([[0, 0], [0, 1.4], [-1.1, 0], [-1.0, 0], [0, 1.5]])
    ph = rh.editor_from_arrays(ns, mass, den, pos, vel)
    out.update(C_mass=mass, C_den=den, C_pos0=pos, C_vel0=vel)
    out["C_check0"] = np.array(ph.check_collision())
    d1 = ph.inelastic_collision()
    out["C_del1"] = np.array(d1)
    out.update(snap(ph, "C_m1_"))
    out["C_check1"] = np.array(ph.check_collision())
    save("editor_clouds", **out)


def gen_pair_kernels(ns):
    """Direct calls of the jitted per-body kernels on seeded clouds."""
    rng = np.random.default_rng(20240612)
    pe = ns.phys_editor
    out = {}
    n = 384
    mass, den, pos, vel = cloud(rng, n, 3.0e3, central_mass=1e4)
    coi = rng.uniform(0, 400, n) * (rng.uniform(0, 1, n) < 0.5)
    coi[0] = 0
    order = np.argsort(mass)[-1::-1]
    parents = np.array(list(map(pe.find_parent_one, list(range(n)), repeat(order), repeat(pos),
                                repeat(coi))))
    out.update(fp_mass=mass, fp_pos=pos, fp_coi=coi, fp_order=order, fp_parents=parents)
    rel_pos = pos[parents] - pos
    acc = np.array(list(map(pe.gravity_one, list(range(n)), parents, repeat(mass), rel_pos,
                            repeat(0.75))))
    out.update(g1_acc=acc, g1_gc=0.75)
    vals = list(map(pe.kepler_basic_one, list(range(n)), parents, repeat(mass), repeat(pos),
                    repeat(vel), repeat(0.75), repeat(0.7)))
    f, ev, a, b, w, c = list(map(np.array, zip(*vals)))
    out.update(kb_vel=vel, kb_focus=f, kb_ecc_v=ev, kb_semi_major=a, kb_semi_minor=b,
               kb_periapsis_arg=w, kb_coi=c)
    # collision predicate incl. exact-touch and one-ulp cases
    m = 256
    cm, cden, cpos, cvel = cloud(rng, m, 9.0e4, central_mass=50.0)
    cpos += 7.0e4                                   # off-origin: no exact zeros
    rad = rng.uniform(2.0, 9.0, m)
    cases = []
    for trial in range(24):
        p = cpos.copy()
        r = rad.copy()
        for extra in range(trial if trial < 4 else 0):   # plain overlaps (several at once)
            i, j = sorted(rng.choice(m, 2, replace=False))
            p[j] = p[i] + 0.5 * (r[i] + r[j]) * np.array([0.6, 0.8])
        if trial >= 4:   # engineer a near-threshold pair
            i, j = sorted(rng.choice(m, 2, replace=False))
            th = rng.uniform(0, 2 * np.pi)
            dist = r[i] + r[j]
            p[j] = p[i] + dist * np.array([np.cos(th), np.sin(th)])
            if trial % 3 == 0:
                p[j, 0] = np.nextafter(p[j, 0], np.inf)
            if trial % 3 == 1:
                p[j, 0] = np.nextafter(p[j, 0], -np.inf)
        val = list(map(pe.check_collision_one, list(range(m - 1)), repeat(cm), repeat(p),
                       repeat(r)))
        b1, b2 = next((it for it in val if it is not None), (-1, -1))
        cases.append((p, r, b1, b2))
    out.update(cc_mass=cm, cc_pos=np.array([c[0] for c in cases]),
               cc_rad=np.array([c[1] for c in cases]),
               cc_del=np.array([c[2] for c in cases]), cc_add=np.array([c[3] for c in cases]))
    # simplified_orbit_coi through the reference object
    ph = rh.editor_from_arrays(ns, mass[:96], den[:96], pos[:96], vel[:96])
    ph.pos, ph.vel = pos[:96].copy(), vel[:96].copy()
    ph.simplified_orbit_coi()
    out.update(soc_coi=ph.coi.copy(), soc_largest=int(ph.largest))
    save("pair_kernels", **out)


def gen_allpairs(ns):
    """All-pairs model: njit driver that calls the reference's own jitted
    gravity_one per pair + documented v+=a; P+=v."""
    from numba import njit
    g1 = ns.phys_editor.gravity_one

    @njit
    def accel(mass, pos, gc):
        n = mass.shape[0]
        acc = np.zeros((n, 2))
        m2 = np.empty(2)
        for i in range(n):
            for j in range(n):
                if j != i:
                    # body index 1 => "not root"; mass looked up via a 2-vector
                    m2[0] = mass[j]
                    m2[1] = mass[i]
                    a = g1(1, 0, m2, pos[j] - pos[i], gc)
                    acc[i, 0] += a[0]
                    acc[i, 1] += a[1]
        return acc

    rng = np.random.default_rng(20240613)
    n = 256
    mass, den, pos, vel = cloud(rng, n, 5.0e3, central_mass=1e4)
    gc = 1.0
    out = dict(mass=mass, pos0=pos, vel0=vel, gc=gc, acc0=accel(mass, pos, gc))
    p, v = pos.copy(), vel.copy()
    for s in range(1, 17):
        v += accel(mass, p, gc)
        p += v
        if s in (1, 16):
            out["pos%d" % s] = p.copy()
            out["vel%d" % s] = v.copy()
    save("allpairs_small", **out)


def gen_kepler_solvers(ns):
    rng = np.random.default_rng(20240614)
    ske = ns.enhanced_kepler_solver.solve_kepler_ell
    nrh = ns.phys_shared.newton_root_kepler_hyp
    ecc = np.concatenate([rng.uniform(0, 0.999, 600), rng.uniform(0.9901, 0.99999, 100),
                          [0.0, 0.1, 0.5, 0.9, 0.0167, 0.7, 0.995, 0.99, 0.9900001]])
    ma = np.concatenate([rng.uniform(0, 2 * np.pi, 500), rng.uniform(-1, 9, 100),
                         rng.uniform(0, 0.0045, 60), 2 * np.pi - rng.uniform(0, 0.0045, 40),
                         [0.0, 1.0, 3.0, 0.1, 4.5, 6.0, 0.001, 0.004, np.pi]])
    ell = np.array([ske(e, m, 1e-10) for e, m in zip(ecc, ma)])
    ell_tol6 = np.array([ske(e, m, 1e-6) for e, m in zip(ecc, ma)])
    hecc = np.concatenate([rng.uniform(1.0001, 4.0, 400), [1.5, 3.0, 1.01, 1.0]])
    hma = np.concatenate([rng.uniform(-6, 6, 400), [2.0, -4.0, 0.5, 0.3]])
    hguess = np.concatenate([hma[:200], rng.uniform(-3, 3, 200), [2.0, 0.0, 0.5, 0.3]])
    hyp = np.array([nrh(e, m, g) for e, m, g in zip(hecc, hma, hguess)])
    save("kepler_solvers", ecc=ecc, ma=ma, ell=ell, ell_tol6=ell_tol6, hecc=hecc, hma=hma,
         hguess=hguess, hyp=hyp)


def gen_game(ns, conf, bd, bod):
    """Game mode: Solar System -> convert.to_kepler -> phys_body (C1,C5,C6,C8);
    synthetic vessels through phys_vessel.Physics.move (C4)."""
    import copy
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    P = 32
    korb = convert.to_kepler(bd["mass"], copy.deepcopy(bod), conf["gc"], conf["coi_coef"])
    pb = ns.phys_body.Physics()
    rh.set_curve_points(ns, pb, P)
    pb.load(conf, copy.deepcopy(bd), copy.deepcopy(korb))
    out = dict(P=P, gc=conf["gc"], coi_coef=conf["coi_coef"], mass=bd["mass"].copy(),
               a=korb["a"].copy(), ecc=korb["ecc"].copy(), pea=korb["pe_arg"].copy(),
               ma0=korb["ma"].copy(), ref=korb["ref"].copy(), dr=korb["dir"].copy())
    body_data, body_orb, pos, ma, ea, cm = pb.initial(1)
    for k in ("b", "f", "coi", "pe_d", "ap_d", "n", "per", "u"):
        out["orb_" + k] = np.array(body_orb[k])
    out.update(i_pos=pos.copy(), i_ma=ma.copy(), i_ea=ea.copy(), i_curves=pb.curves.copy(),
               i_curves_mov=cm.copy())
    for k in range(1, 1001):
        pos, ma, ea = pb.move(3 if k % 2 else 1)
        if k in (1, 1000):
            out.update({"m%d_pos" % k: pos.copy(), "m%d_ma" % k: ma.copy(),
                        "m%d_ea" % k: ea.copy(), "m%d_curves_mov" % k: pb.curve_move().copy()})
    body_pos_final = pos.copy()

    # synthetic vessels around the solar-system bodies (mixed elliptic/hyperbolic)
    rng = np.random.default_rng(20240615)
    nv = 200
    pv = ns.phys_vessel.Physics()
    rh.set_curve_points(ns, pv, P)
    v_ref = rng.integers(0, len(bd["mass"]), nv)
    hypm = rng.uniform(0, 1, nv) < 0.3
    v_ecc = np.where(hypm, rng.uniform(1.02, 3.0, nv), rng.uniform(0.0, 0.97, nv))
    v_ecc[:6] = [0.0, 0.995, 0.999, 1.0, 0.9905, 0.5]
    hypm = v_ecc >= 1
    v_a = rng.uniform(50, 4000, nv) * np.where(v_ecc > 1, -1, 1)
    v_pea = rng.uniform(0, 2 * np.pi, nv)
    v_ma = np.where(hypm, rng.uniform(-3, 3, nv), rng.uniform(0, 2 * np.pi, nv))
    v_ma[1:3] = [0.001, 2 * np.pi - 0.002]     # ENRKE quirk region (F8b)
    v_dr = rng.choice([-1.0, 1.0], nv)
    vals = list(map(ns.phys_vessel.calc_orb_one, v_ref, repeat(bd["mass"]), repeat(conf["gc"]),
                    v_a, v_ecc))
    vb, vf, vpe, vap, vper, vn, vu = list(map(np.array, zip(*vals)))
    pv.names = np.array(["v%d" % i for i in range(nv)])
    pv.a, pv.ecc, pv.pea, pv.ma, pv.ref, pv.dr = v_a, v_ecc, v_pea, v_ma.copy(), v_ref, v_dr
    pv.b, pv.f, pv.n, pv.u = vb, vf, vn, vu
    pv.pos = np.zeros((nv, 2))
    pv.ea = np.where(hypm, v_ma, 0.0)
    pv.curves = np.zeros((nv, P, 2))
    out.update(v_ref=v_ref, v_ecc=v_ecc, v_a=v_a, v_pea=v_pea, v_ma0=v_ma, v_dr=v_dr,
               v_ea0=pv.ea.copy(), v_b=vb, v_f=vf, v_n=vn, v_u=vu, v_pe_d=vpe, v_ap_d=vap,
               v_per=vper, v_body_pos=body_pos_final)
    for k in range(1, 51):
        vpos, vma = pv.move(2, body_pos_final, ma, ea)
        if k in (1, 50):
            out.update({"v%d_pos" % k: vpos.copy(), "v%d_ma" % k: vma.copy(),
                        "v%d_ea" % k: pv.ea.copy()})
    for v in range(nv):
        pv.curve(v)
    out.update(v_curves=pv.curves.copy(), v_curves_mov=pv.curve_move().copy())
    save("game_mode", **out)


def gen_culling(ns, conf, bd, bod):
    """N3: phys_shared.culling (jitted) on random clouds, phys_body / phys_vessel culling on the
    Solar System game state + synthetic vessels."""
    import copy
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    rng = np.random.default_rng(20240616)
    out = {}
    n = 500
    coords = rng.uniform(-2000, 2000, (n, 2))
    radius = rng.uniform(0, 60, n)
    cases = []
    for k in range(12):
        c = rng.uniform(-1500, 1500, 2)
        half = rng.uniform(50, 1500, 2)
        zoom = float(rng.choice([0.05, 0.3, 1.0, 4.0]))
        sb = np.array([[c[0] - half[0], c[1] + half[1]], [c[0] + half[0], c[1] - half[1]]])
        if k == 0:   # touch a bound exactly
            sb[0, 0] = coords[7, 0] + radius[7] + 2 / zoom
        cases.append((sb, zoom, ns.phys_shared.culling(coords, radius, sb, zoom)))
    out.update(c_coords=coords, c_radius=radius, c_sb=np.array([c[0] for c in cases]),
               c_zoom=np.array([c[1] for c in cases]), c_mask=np.array([c[2] for c in cases]))
    # body culling on the Solar System (game state after initial)
    korb = convert.to_kepler(bd["mass"], copy.deepcopy(bod), conf["gc"], conf["coi_coef"])
    pb = ns.phys_body.Physics()
    rh.set_curve_points(ns, pb, 16)
    pb.load(conf, copy.deepcopy(bd), copy.deepcopy(korb))
    body_data, body_orb, pos, ma, ea, cm = pb.initial(1)
    out.update(b_pos=pos.copy(), b_radius=pb.radius.copy(), b_atm_h=pb.atm_h.copy(),
               b_coi=pb.coi.copy(), b_ref=np.array(pb.ref), b_a=pb.a.copy(),
               b_screen_x=float(pb.screen_x))
    nv = 300
    pv = ns.phys_vessel.Physics()
    pv.names = np.array(["v%d" % i for i in range(nv)])
    pv.pos = pos[rng.integers(0, len(pos), nv)] + rng.uniform(-300, 300, (nv, 2))
    pv.ref = rng.integers(0, len(pos), nv)
    pv.pe_d = rng.uniform(10, 40000, nv)
    pv.body_dr = np.ones(len(pos))
    pv.body_coi = pb.coi.copy()
    out.update(v_pos=pv.pos.copy(), v_ref=pv.ref.copy(), v_pe_d=pv.pe_d.copy())
    bcases = []
    for k in range(8):
        centre = pos[int(rng.integers(0, len(pos)))] if k % 2 else np.zeros(2)
        half = np.array([683.0, 384.0]) / [0.02, 0.1, 0.5, 2.0][k % 4]
        zoom = [0.02, 0.1, 0.5, 2.0][k % 4]
        sb = np.array([[centre[0] - half[0], centre[1] + half[1]],
                       [centre[0] + half[0], centre[1] - half[1]]])
        vb, vc, vo = pb.culling(sb, zoom)
        vv, vvo = pv.culling(sb, zoom, vc)
        bcases.append((sb, zoom, vb, vc, vo, np.array(vv), np.array(vvo)))
    out["b_sb"] = np.array([c[0] for c in bcases])
    out["b_zoom"] = np.array([c[1] for c in bcases])
    for k, c in enumerate(bcases):
        out.update({"b%d_bodies" % k: c[2], "b%d_coi" % k: c[3], "b%d_orbits" % k: c[4],
                    "b%d_vessels" % k: c[5], "b%d_vorbits" % k: c[6]})
    save("culling", **out)


def gen_body_data(ns, conf, bd, bod):
    """N4: reference Physics.body / temp_color / classify on the Solar System and on a synthetic
    population spanning dwarf planets to black holes."""
    import copy
    rng = np.random.default_rng(20240617)
    out = {}
    for tag in ("solar", "synth"):
        if tag == "solar":
            body_data, orb = copy.deepcopy(bd), copy.deepcopy(bod)
        else:
            n = 400
            mass = 10 ** rng.uniform(-2, 7.5, n)
            mass[0] = 10 ** 8
            den_arr = 10 ** rng.uniform(2, 5, n)
            den_arr[1:9] = 10 ** rng.uniform(14, 22, 8)     # compact objects -> black holes
            mass[1:9] = 10 ** rng.uniform(5, 8, 8)
            body_data = {
                "name": np.array(["b%d" % i for i in range(n)]), "mass": mass,
                "den": den_arr,
                "color": rng.integers(0, 256, (n, 3)),
                "atm_pres0": np.where(rng.uniform(0, 1, n) < 0.5, rng.uniform(0.1, 90, n), 0.0),
                "atm_scale_h": np.where(rng.uniform(0, 1, n) < 0.7, rng.uniform(0.005, 0.02, n), 0.0),
                "atm_den0": np.where(rng.uniform(0, 1, n) < 0.7, rng.uniform(0.5, 70, n), 0.0)}
            r = rng.uniform(1e3, 1e6, n)
            th = rng.uniform(0, 2 * np.pi, n)
            orb = {"kepler": False, "pos": np.stack([r * np.cos(th), r * np.sin(th)], axis=1),
                   "vel": rng.uniform(-1, 1, (n, 2))}
            orb["pos"][0] = 0
        ph = ns.phys_editor.Physics()
        ph.load_system(conf, body_data, orb)
        color = ph.temp_color()
        ph.classify()
        out.update({tag + "_mass": np.array(ph.mass), tag + "_den": np.array(ph.den),
                    tag + "_rad": np.array(ph.rad),
                    tag + "_atm_pres0": np.array(ph.atm_pres0),
                    tag + "_atm_scale_h": np.array(ph.atm_scale_h),
                    tag + "_atm_den0": np.array(ph.atm_den0),
                    tag + "_base_color": np.array(ph.base_color),
                    tag + "_surf_grav": np.array(ph.surf_grav), tag + "_atm_h": np.array(ph.atm_h),
                    tag + "_luminosity": np.array(ph.luminosity), tag + "_temp": np.array(ph.temp),
                    tag + "_rad_sc": np.array(ph.rad_sc), tag + "_color": np.array(color),
                    tag + "_types": np.array(ph.types),
                    tag + "_stellar_class": np.array(ph.stellar_class)})
    out.update(gc=conf["gc"], rad_mult=conf["rad_mult"], mass_thermal_mult=conf["mass_thermal_mult"],
               min_mass=conf["min_planet_mass"])
    save("body_data", **out)


def gen_editor_edits(ns):
    """Boundary mutators of phys_editor.Physics: add_body (light, and heavier than
    the root -> re-root), set_body_mass (incl. root change), set_body_pos/vel/den, del_body
    (ordinary and root), each followed by a few frames; snapshots after every step."""
    rng = np.random.default_rng(20240618)
    n = 24
    mass, den, pos, vel = cloud(rng, n, 3.0e4, central_mass=2.0e4)
    mass[1:5] = rng.uniform(100, 900, 4)
    den[:] = 3.0e7
    ph = rh.editor_from_arrays(ns, mass, den, pos, vel)
    out = dict(mass=mass, den=den, pos0=pos, vel0=vel)
    steps = []

    def snap_step(tag):
        ph.kepler_basic(); ph.temp_color(); ph.classify()
        for _ in range(3):
            rh.editor_frame(ph)
        out.update(snap(ph, "s%d_" % len(steps)))
        out["s%d_largest" % len(steps)] = int(ph.largest)
        steps.append(tag)

    def new_body(m, p, v):
        return {"name": "new", "mass": m, "density": 3.0e7, "position": np.array(p, float),
                "velocity": np.array(v, float), "color": (1, 2, 3), "atm_pres0": 0.0,
                "atm_scale_h": 0.0, "atm_den0": 0.0}

    snap_step("load")
    ph.add_body(new_body(3.0, [9000.0, 2000.0], [0.1, 1.2])); snap_step("add_light")
    ph.set_body_mass(7, 1500.0); snap_step("set_mass")
    ph.set_body_pos(9, np.array([-4000.0, 7000.0])); snap_step("set_pos")
    ph.set_body_vel(9, np.array([0.9, 0.4])); snap_step("set_vel")
    ph.set_body_den(3, 0.01); snap_step("set_den")
    ph.del_body(5); snap_step("del_ordinary")
    ph.add_body(new_body(9.0e4, [15000.0, -8000.0], [0.0, 0.3])); snap_step("add_new_root")
    ph.set_body_mass(4, 5.0e5); snap_step("set_mass_new_root")
    ph.del_body(0); snap_step("del_root")
    ph.kepler_inverse(6, 0.3, 40.0, 2000.0, 75.0, 0.0, 0.0, 1.0); snap_step("kepler_inverse_ma")
    ph.kepler_inverse(8, 0.5, 200.0, 1500.0, 10.0, 120.0, 5200.0, -1.0)
    snap_step("kepler_inverse_ta_apd")
    out["steps"] = np.array(steps)
    save("editor_edits", **out)


def gen_convert(ns, conf, bd, bod):
    """N1: convert.to_kepler / to_newton on the Solar System and on hierarchical clouds whose
    index order differs from their mass order (the reference's index quirks matter there)."""
    import copy
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    rng = np.random.default_rng(20240619)
    out = {}
    cases = {"solar": (bd["mass"].copy(), bod["pos"].copy(), bod["vel"].copy())}
    for tag, n in (("c60", 60), ("c1500", 1500)):
        mass, den, pos, vel = cloud(rng, n, 3.0e4, central_mass=2.0e4)
        npl = 6 if n == 60 else 40
        mass[1:1 + npl] = rng.uniform(50, 900, npl)
        for k in range(1 + npl, 1 + 3 * npl):          # moons inside planet COIs
            p = 1 + (k % npl)
            d = rng.uniform(30, 90)
            th = rng.uniform(0, 2 * np.pi)
            pos[k] = pos[p] + d * np.array([np.cos(th), np.sin(th)])
            vel[k] = vel[p] + np.sqrt(mass[p] / d) * np.array([-np.sin(th), np.cos(th)])
            mass[k] = rng.uniform(0.01, 0.4)
        perm = np.concatenate(([0], 1 + rng.permutation(n - 1)))   # scramble all but the root
        cases[tag] = (mass[perm], pos[perm], vel[perm])
    for tag, (mass, pos, vel) in cases.items():
        k = convert.to_kepler(mass, {"pos": pos.copy(), "vel": vel.copy()}, conf["gc"],
                              conf["coi_coef"])
        out.update({tag + "_mass": mass, tag + "_pos": pos, tag + "_vel": vel})
        for key in ("a", "ecc", "pe_arg", "ma", "ref", "dir"):
            out[tag + "_k_" + key] = np.array(k[key])
        if np.all(k["ecc"] < 1):
            nw = convert.to_newton(mass, copy.deepcopy(k), conf["gc"], conf["coi_coef"])
            out[tag + "_n_pos"] = nw["pos"]
            out[tag + "_n_vel"] = nw["vel"]
    # to_newton on synthetic elliptic elements with arbitrary refs (incl. ref >= body)
    n = 300
    mass = rng.uniform(0.5, 50.0, n)
    mass[0] = 2.0e4
    kn = {"kepler": True, "a": rng.uniform(200, 2.0e4, n), "ecc": rng.uniform(0, 0.9, n),
          "pe_arg": rng.uniform(0, 2 * np.pi, n), "ma": rng.uniform(0, 2 * np.pi, n),
          "ref": rng.integers(0, n, n), "dir": rng.choice([-1.0, 1.0], n)}
    kn["ecc"][[5, 17]] = 0.0
    kn["ref"][:40] = 0
    kn["ref"][0] = 0
    nw = convert.to_newton(mass, copy.deepcopy(kn), conf["gc"], conf["coi_coef"])
    out.update(kn_mass=mass, kn_n_pos=nw["pos"], kn_n_vel=nw["vel"])
    for key in ("a", "ecc", "pe_arg", "ma", "ref", "dir"):
        out["kn_k_" + key] = np.array(kn[key])
    out.update(gc=conf["gc"], coi_coef=conf["coi_coef"])
    save("convert", **out)


def gen_intersect(ns):
    """N2 numeric kernels: solve_quartic, ell_hyp_intersect_circle, predict_enter_coi (jitted)."""
    sys.path.insert(0, rh.REFERENCE_ROOT)
    os.chdir(rh._WORKDIR)
    from volatilespace.physics import orbit_intersect, quartic_solver
    sys.path.remove(rh.REFERENCE_ROOT)
    rng = np.random.default_rng(20240620)
    out = {}
    coef = rng.uniform(-10, 10, (300, 5))
    coef[:, 0] = rng.uniform(0.5, 5, 300) * rng.choice([-1, 1], 300)
    coef[:40, 1] = 0.0
    coef[:40, 3] = 0.0                       # biquadratics
    roots = np.array([quartic_solver.solve_quartic(*c) for c in coef])
    out.update(q_coef=coef, q_roots_re=roots.real, q_roots_im=roots.imag)
    # orbit / circle intersections: ellipses and hyperbolas, circle centred on the focus
    n = 400
    hyp = rng.uniform(0, 1, n) < 0.4
    ecc = np.where(hyp, rng.uniform(1.05, 3.0, n), rng.uniform(0.0, 0.95, n))
    a = rng.uniform(100, 5000, n) * np.where(hyp, -1, 1)
    f = a * ecc
    b = np.sqrt(np.abs(f**2 - a**2))
    cx = f.copy()
    cy = np.where(rng.uniform(0, 1, n) < 0.8, 0.0, rng.uniform(-200, 200, n))
    r = np.abs(a) * rng.uniform(0.05, 2.5, n)
    eas = np.full((n, 4), np.nan)
    cnt = np.zeros(n, dtype=np.int64)
    for i in range(n):
        e = orbit_intersect.ell_hyp_intersect_circle(a[i], b[i], ecc[i], cx[i], cy[i], r[i])
        if not (len(e) == 1 and np.isnan(e[0])):
            cnt[i] = len(e)
            eas[i, :len(e)] = e
    out.update(i_a=a, i_b=b, i_ecc=ecc, i_cx=cx, i_cy=cy, i_r=r, i_ea=eas, i_cnt=cnt)
    # predict_enter_coi: vessels and bodies around the same reference
    m = 120
    u = 1.0e4
    res = np.zeros((m, 5))
    vds, bds = np.zeros((m, 10)), np.zeros((m, 10))
    for i in range(m):
        va = rng.uniform(2000, 9000)
        vecc = rng.uniform(0.0, 0.7)
        vf = va * vecc
        vb = np.sqrt(abs(vf**2 - va**2))
        vper = 2 * np.pi * np.sqrt(va**3 / u)
        vn = 2 * np.pi / vper
        ba = rng.uniform(3000, 8000)
        becc = rng.uniform(0.0, 0.3)
        bf = ba * becc
        bb = np.sqrt(abs(bf**2 - ba**2))
        bn = np.sqrt(u / ba**3)
        vd = (va, vb, vf, vecc, rng.uniform(0, 2 * np.pi), 0.0, rng.uniform(0, 2 * np.pi),
              vper if i % 10 else 0.0, vn, float(rng.choice([-1, 1])))
        vd = vd[:5] + (ns.enhanced_kepler_solver.solve_kepler_ell(vecc, vd[4], 1e-10),) + vd[6:]
        bma = rng.uniform(0, 2 * np.pi)
        bd = (ba, bb, bf, becc, bma, ns.enhanced_kepler_solver.solve_kepler_ell(becc, bma, 1e-10),
              rng.uniform(0, 2 * np.pi), bn, float(rng.choice([-1, 1])),
              rng.uniform(600, 3000) if i % 7 else 0.0)
        res[i] = orbit_intersect.predict_enter_coi(vd, bd, 1e-5, 300)
        vds[i], bds[i] = vd, bd
    out.update(p_vd=vds, p_bd=bds, p_out=res)
    save("intersect", **out)


def gen_vessel_points(ns, conf, bd, bod):
    """N2 bookkeeping: phys_vessel.Physics.points via load() + initial() of the live reference on
    the Solar System with synthetic vessels (elliptic and hyperbolic, around the Sun, planets with
    moons and moons), then single-vessel points() calls with the entered_coi / left_coi /
    left_coi_prev_ref / only_enter_coi variants."""
    import copy
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    P = 8
    korb = convert.to_kepler(bd["mass"], copy.deepcopy(bod), conf["gc"], conf["coi_coef"])
    pb = ns.phys_body.Physics()
    rh.set_curve_points(ns, pb, P)
    pb.load(conf, copy.deepcopy(bd), copy.deepcopy(korb))
    body_data, body_orb, bpos, bma, bea, _ = pb.initial(1)
    nb = len(bd["mass"])
    rng = np.random.default_rng(20240622)
    nv = 160
    coi = np.array(body_orb["coi"], dtype=float)
    size = np.array(body_data["radius"], dtype=float)
    bref = np.array(body_orb["ref"])
    ba = np.array(body_orb["a"], dtype=float)
    parents = np.unique(bref[1:])                  # bodies that have satellites
    v_ref = np.where(rng.uniform(0, 1, nv) < 0.7, rng.choice(parents, nv), rng.integers(0, nb, nv))
    v_ref[:40] = 0
    hypm = rng.uniform(0, 1, nv) < 0.35
    v_ecc = np.where(hypm, rng.uniform(1.05, 2.5, nv), rng.uniform(0.0, 0.9, nv))
    v_ecc[40:44] = [0.0, 0.3, 1.6, 0.95]
    hypm = v_ecc >= 1
    # scale of each orbit: between the parent's surface and its COI (the Sun: out to the planets)
    span = np.where(v_ref == 0, ba.max() * 1.2, coi[v_ref])
    lo = size[v_ref] * 0.8
    pe = lo + (span * 0.9 - lo) * rng.uniform(0, 1, nv) ** 2
    v_a = np.where(hypm, -pe / (v_ecc - 1), pe / (1 - np.minimum(v_ecc, 0.9999)))
    v_pea = rng.uniform(0, 2 * np.pi, nv)
    v_ma = np.where(hypm, rng.uniform(-1.5, 0.5, nv), rng.uniform(0, 2 * np.pi, nv))
    v_dr = rng.choice([-1.0, 1.0], nv)
    vessel_data = {"name": np.array(["v%d" % i for i in range(nv)]), "mass": np.ones(nv),
                   "rot_angle": np.zeros(nv), "rot_acc": np.zeros(nv)}
    vessel_orb = {"a": v_a.copy(), "ecc": v_ecc.copy(), "pe_arg": v_pea.copy(), "ma": v_ma.copy(),
                  "ref": v_ref.copy(), "dir": v_dr.copy()}
    pv = ns.phys_vessel.Physics()
    rh.set_curve_points(ns, pv, P)
    pv.load(conf, body_data, body_orb, vessel_data, vessel_orb)
    pv.initial(1, bpos, bma, bea)

    def vessel12():
        return np.stack([pv.a, pv.b, pv.f, pv.ecc, pv.ma, pv.ea, pv.pea, pv.n, pv.dr, pv.pe_d,
                         pv.ap_d, pv.period], axis=1).astype(np.float64)

    body12 = np.stack([pv.body_a, pv.body_b, pv.body_f, pv.body_ecc, pv.body_ma, pv.body_ea,
                       pv.body_pea, pv.body_n, pv.body_dr, pv.body_coi, pv.body_size,
                       pv.body_atm], axis=1).astype(np.float64)
    out = dict(vessel12=vessel12(), v_ref=np.array(pv.ref, dtype=np.int64), body12=body12,
               body_ref=np.array(pv.body_ref, dtype=np.int64),
               steps_limit=pv.predict_coi_limit,
               body_impact=pv.body_impact.copy(), body_enter_atm=pv.body_enter_atm.copy(),
               coi_leave=pv.coi_leave.copy(), coi_enter=pv.coi_enter.copy())
    # single-vessel variants (attributes the reference's cross_coi sets before calling points)
    var = []
    cand = [v for v in range(nv) if not np.isnan(pv.coi_enter[v, 0])][:6] + list(range(44, 52))
    for k, v in enumerate(cand):
        mode = k % 4
        pv.entered_coi = v if mode == 0 else None
        pv.left_coi = v if mode in (1, 2) else None
        prev = None
        if mode == 2:
            sib = np.where(np.array(pv.body_ref) == pv.ref[v])[0]
            sib = sib[sib != 0]
            prev = int(sib[0]) if len(sib) else None
        pv.left_coi_prev_ref = prev
        before = pv.coi_leave.copy()
        pv.points(v, only_enter_coi=(mode == 3))
        var.append([v, -1 if pv.entered_coi is None else pv.entered_coi,
                    -1 if pv.left_coi is None else pv.left_coi, -1 if prev is None else prev,
                    int(mode == 3)])
        out["var%d_coi_leave_in" % k] = before
        out["var%d_rows" % k] = np.concatenate([pv.body_impact[v], pv.body_enter_atm[v],
                                                pv.coi_leave[v], pv.coi_enter[v]])
    pv.entered_coi = pv.left_coi = pv.left_coi_prev_ref = None
    out["variants"] = np.array(var, dtype=np.int64)
    save("vessel_points", **out)
    print("  vessel_points: impact %d  atm %d  leave %d  enter %d of %d vessels" % (
        np.sum(~np.isnan(pv.body_impact[:, 0])), np.sum(~np.isnan(pv.body_enter_atm[:, 0])),
        np.sum(~np.isnan(pv.coi_leave[:, 0])), np.sum(~np.isnan(pv.coi_enter[:, 0])), nv))


def gen_vessel_events(ns, conf, bd, bod):
    """N2 per-tick vessel events: the physics loop of game.py:1165-1265 (check_warp,
    predict_enter_coi_service, body/vessel move, cross_coi, change_vessel) run on the live
    reference for the Solar System + the synthetic vessels of `vessel_points`; function-level
    snapshots (inputs and outputs of each call) at the ticks where something happens and at a few
    quiet ones."""
    import copy
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    g = np.load(os.path.join(OUT, "vessel_points.npz"))
    P = 8
    korb = convert.to_kepler(bd["mass"], copy.deepcopy(bod), conf["gc"], conf["coi_coef"])
    pb = ns.phys_body.Physics()
    rh.set_curve_points(ns, pb, P)
    pb.load(conf, copy.deepcopy(bd), copy.deepcopy(korb))
    body_data, body_orb, bpos, bma, bea, _ = pb.initial(1)
    v12 = g["vessel12"]
    # leave out the high-eccentricity ellipses (their COI-entry march is chaotic in the reference,
    # see DESIGN.md) and the four vessels whose march takes tens of seconds per call
    keep = np.flatnonzero(~((v12[:, 3] >= 0.8) & (v12[:, 3] < 1)))
    keep = keep[~np.isin(keep, [2, 5])]
    nv = len(keep)
    vessel_data = {"name": np.array(["v%d" % i for i in keep]), "mass": np.ones(nv),
                   "rot_angle": np.zeros(nv), "rot_acc": np.zeros(nv)}
    ma0 = v12[keep, 4].copy()
    vessel_orb = {"a": v12[keep, 0].copy(), "ecc": v12[keep, 3].copy(), "pe_arg": v12[keep, 6].copy(),
                  "ma": ma0, "ref": g["v_ref"][keep].copy(), "dir": v12[keep, 8].copy()}
    pv = ns.phys_vessel.Physics()
    rh.set_curve_points(ns, pv, P)
    pv.load(conf, body_data, body_orb, vessel_data, vessel_orb)
    # initial() moves once by `warp`; start from the fixture's pre-move anomalies
    hyp = np.where(pv.ecc > 1, -1, 1)
    pv.ma = pv.ma - 0.0
    pv.initial(1, bpos, bma, bea)
    out = dict(keep=keep, gc=conf["gc"], body_mass=np.array(pv.body_mass, dtype=float),
               body_ref=np.array(pv.body_ref, dtype=np.int64),
               body_static=np.stack([pv.body_a, pv.body_b, pv.body_f, pv.body_ecc, pv.body_pea,
                                     pv.body_n, pv.body_dr, pv.body_coi, pv.body_size,
                                     pv.body_atm], axis=1).astype(float),
               steps_limit=pv.predict_coi_limit, P=P)

    def vstate():
        return np.stack([pv.a, pv.b, pv.f, pv.ecc, pv.pea, pv.n, pv.dr, pv.ma, pv.ea, pv.prev_ea,
                         pv.u, pv.pe_d, pv.ap_d, pv.period], axis=1).astype(float)

    def pts():
        return np.concatenate([pv.body_impact, pv.body_enter_atm, pv.coi_leave, pv.coi_enter], axis=1)

    def hold_mask():
        m = np.zeros(nv, dtype=np.uint8)
        m[np.asarray(pv.physical_hold, dtype=int)] = 1
        return m

    opt = lambda x: -1 if x is None else int(x)
    service_calls = []
    orig_points = pv.points

    def logged_points(vessel, only_enter_coi=False):
        if only_enter_coi:
            service_calls.append(vessel)
        return orig_points(vessel, only_enter_coi)

    pv.points = logged_points
    cw, sv, cx, cv = [], [], [], []
    warp_cycle = [1, 5, 25, 100]
    T = 6000
    phase_len = 250
    cap = dict(cw=24, sv=16, cx=28, cv=16)          # fixture size: ~20 KB per snapshot
    per_phase = dict(cw=2, sv=1, cx=2, cv=1)
    seen = {}

    def room(kind, lst, tick):
        key = (kind, tick // phase_len)
        if len(lst) >= cap[kind] or seen.get(key, 0) >= per_phase[kind]:
            return False
        seen[key] = seen.get(key, 0) + 1
        return True

    for tick in range(T):
        warp = warp_cycle[(tick // phase_len) % len(warp_cycle)]
        # --- check_warp
        pre = (vstate(), pts(), hold_mask())
        situation, alarm, nhold = pv.check_warp(warp)
        if situation is not None or tick % 1500 == 0 or not np.array_equal(pre[2], hold_mask()):
            if room("cw", cw, tick):
                k = len(cw)
                out.update({"cw%d_v" % k: pre[0], "cw%d_p" % k: pre[1], "cw%d_hold_in" % k: pre[2],
                            "cw%d_hold_out" % k: hold_mask(),
                            "cw%d_out" % k: np.array([opt(situation), opt(alarm), nhold, warp])})
                cw.append(tick)
        # --- predict_enter_coi_service
        del service_calls[:]
        pre = (vstate(), pts())
        flags_in = np.array([opt(pv.entered_coi), opt(pv.left_coi), opt(pv.left_coi_prev_ref)])
        pv.predict_enter_coi_service()
        if (service_calls or tick % 1500 == 0) and room("sv", sv, tick):
            k = len(sv)
            calls = np.array(sorted(set(service_calls)), dtype=np.int64)
            out.update({"sv%d_v" % k: pre[0], "sv%d_p" % k: pre[1], "sv%d_flags" % k: flags_in,
                        "sv%d_calls" % k: calls, "sv%d_bma" % k: np.array(pv.body_ma, dtype=float),
                        "sv%d_bea" % k: np.array(pv.body_ea, dtype=float),
                        "sv%d_ref" % k: np.array(pv.ref, dtype=np.int64),
                        "sv%d_p_out_rows" % k: pts()[calls]})
            sv.append(tick)
        # --- move
        bpos, bma, bea = pb.move(warp)
        pv.move(warp, bpos, bma, bea)
        # --- cross_coi
        pre = (vstate(), pts(), np.array(pv.ref, dtype=np.int64), np.array(bpos, dtype=float),
               np.array([opt(pv.entered_coi), opt(pv.left_coi), opt(pv.left_coi_prev_ref)]))
        changed = pv.cross_coi()
        if (changed is not None or tick % 1500 == 750) and room("cx", cx, tick):
            k = len(cx)
            out.update({"cx%d_v" % k: pre[0], "cx%d_p" % k: pre[1], "cx%d_ref" % k: pre[2],
                        "cx%d_bpos" % k: pre[3], "cx%d_flags_in" % k: pre[4],
                        "cx%d_flags_out" % k: np.array([opt(pv.entered_coi), opt(pv.left_coi),
                                                        opt(pv.left_coi_prev_ref)]),
                        "cx%d_ret" % k: np.array([opt(changed)]),
                        "cx%d_row_out" % k: vstate()[max(opt(changed), 0)],
                        "cx%d_ref_out" % k: np.array([pv.ref[max(opt(changed), 0)]], dtype=np.int64)})
            cx.append(tick)
        # --- change_vessel
        if changed is not None:
            pre = (vstate(), pts(), np.array(pv.ref, dtype=np.int64),
                   np.array([opt(pv.entered_coi), opt(pv.left_coi), opt(pv.left_coi_prev_ref)]),
                   np.array(pv.body_ma, dtype=float), np.array(pv.body_ea, dtype=float))
            pv.change_vessel(changed)
            if room("cv", cv, tick):
                k = len(cv)
                out.update({"cv%d_v" % k: pre[0], "cv%d_p" % k: pre[1], "cv%d_ref" % k: pre[2],
                            "cv%d_flags" % k: pre[3], "cv%d_bma" % k: pre[4], "cv%d_bea" % k: pre[5],
                            "cv%d_vessel" % k: np.array([changed]),
                            "cv%d_row_out" % k: vstate()[changed], "cv%d_p_row_out" % k: pts()[changed],
                            "cv%d_curve" % k: pv.curves[changed].copy()})
                cv.append(tick)
    out.update(n_cw=len(cw), n_sv=len(sv), n_cx=len(cx), n_cv=len(cv))
    save("vessel_events", **out)
    print("  vessel_events: %d vessels, %d check_warp, %d service, %d cross_coi, %d change_vessel "
          "snapshots" % (nv, len(cw), len(sv), len(cx), len(cv)))


def gen_vessel_draw(ns, conf, bd, bod):
    """Drawing-side vessel functions of game mode (phys_vessel.py:635-741 predict_next_orbit,
    :761-774 rotate, :777-825 selected, :828-1075 curve_segments) on the live reference: the
    Solar System + the synthetic vessels of `vessel_points`, snapshots of the inputs and outputs of
    each call at three ticks of the game loop."""
    import copy
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    g = np.load(os.path.join(OUT, "vessel_points.npz"))
    P = 16
    korb = convert.to_kepler(bd["mass"], copy.deepcopy(bod), conf["gc"], conf["coi_coef"])
    pb = ns.phys_body.Physics()
    rh.set_curve_points(ns, pb, P)
    pb.load(conf, copy.deepcopy(bd), copy.deepcopy(korb))
    body_data, body_orb, bpos, bma, bea, _ = pb.initial(1)
    v12 = g["vessel12"]
    keep = np.flatnonzero(~((v12[:, 3] >= 0.8) & (v12[:, 3] < 1)))
    keep = keep[~np.isin(keep, [2, 5])]
    nv = len(keep)
    rng = np.random.default_rng(20240701)
    vessel_data = {"name": np.array(["v%d" % i for i in keep]), "mass": np.ones(nv),
                   "rot_angle": rng.uniform(0, 2 * np.pi, nv), "rot_acc": rng.uniform(0.001, 0.01, nv)}
    vessel_orb = {"a": v12[keep, 0].copy(), "ecc": v12[keep, 3].copy(), "pe_arg": v12[keep, 6].copy(),
                  "ma": v12[keep, 4].copy(), "ref": g["v_ref"][keep].copy(), "dir": v12[keep, 8].copy()}
    pv = ns.phys_vessel.Physics()
    rh.set_curve_points(ns, pv, P)
    pv.load(conf, body_data, body_orb, vessel_data, vessel_orb)
    pv.initial(1, bpos, bma, bea)
    out = dict(keep=keep, gc=conf["gc"], P=P, body_mass=np.array(pv.body_mass, dtype=float),
               body_ref=np.array(pv.body_ref, dtype=np.int64),
               body_static=np.stack([pv.body_a, pv.body_b, pv.body_f, pv.body_ecc, pv.body_pea,
                                     pv.body_n, pv.body_dr, pv.body_coi], axis=1).astype(float),
               rot_acc=np.array(pv.rot_acc, dtype=float))

    def vstate():
        return np.stack([pv.a, pv.b, pv.f, pv.ecc, pv.pea, pv.n, pv.dr, pv.ma, pv.ea, pv.prev_ea,
                         pv.u, pv.pe_d, pv.ap_d, pv.period], axis=1).astype(float)

    def pts():
        return np.concatenate([pv.body_impact, pv.body_enter_atm, pv.coi_leave, pv.coi_enter], axis=1)

    snaps = 0
    cover = {}
    tick = 0
    for stop, warp in ((0, 1), (160, 5), (420, 25)):
        while tick < stop:
            pv.check_warp(warp)
            pv.predict_enter_coi_service()
            bpos, bma, bea = pb.move(warp)
            pv.move(warp, bpos, bma, bea)
            changed = pv.cross_coi()
            if changed is not None:
                pv.change_vessel(changed)
            tick += 1
        k = snaps
        cm = pv.curve_move()
        pv.visible_orbits = list(range(nv))
        out.update({"s%d_v" % k: vstate(), "s%d_p" % k: pts(), "s%d_ref" % k: np.array(pv.ref, dtype=np.int64),
                    "s%d_pos" % k: np.array(pv.pos, dtype=float), "s%d_bpos" % k: np.array(bpos, dtype=float),
                    "s%d_bma" % k: np.array(pv.body_ma, dtype=float), "s%d_curves_mov" % k: np.array(cm, dtype=float)})
        light, dark, first, itype, itime, srange = pv.curve_segments()
        itype = np.array(itype, dtype=np.int64)
        # dark is np.empty garbage where the reference never writes it: no intersection at all
        dark_valid = itype > 0
        for v in range(nv):
            al = np.array([pv.body_impact[v, int(pv.ecc[v] >= 1)], pv.coi_enter[v, 1],
                           pv.coi_leave[v, int(pv.ecc[v] >= 1)]])
            if not dark_valid[v] and not np.all(np.isnan(al)):
                dark_valid[v] = True          # cut by a later COI leave, or the full curve
            key = (int(itype[v]), bool(pv.ecc[v] < 1), bool(pv.dr[v] > 0))
            cover[key] = cover.get(key, 0) + 1
        dark = np.where(dark_valid[:, None, None], dark, np.nan)
        out.update({"s%d_light" % k: light, "s%d_dark" % k: dark, "s%d_dark_valid" % k: dark_valid,
                    "s%d_first" % k: first, "s%d_itype" % k: itype,
                    "s%d_itime" % k: np.array(itime, dtype=float), "s%d_srange" % k: srange})
        # selected(): every vessel the reference's scalar math accepts
        sel = np.full((nv, 13), np.nan)
        for v in range(nv):
            try:
                ta, pe, pe_t, ap, ap_t, dist, so, sh, sv = pv.selected(v)
                sel[v] = [ta, pe[0], pe[1], pe_t, ap[0], ap[1], ap_t, dist, so, sh, sv, 1.0, 0.0]
            except (ValueError, ZeroDivisionError):
                pass
        out["s%d_selected" % k] = sel
        # predict_next_orbit(): vessels with a COI change ahead
        pno_v, pno_curve, pno_sc = [], [], []
        for v in range(nv):
            try:
                r = pv.predict_next_orbit(v)
            except (ValueError, ZeroDivisionError, FloatingPointError):
                continue
            if r[0] is None:
                continue
            curve, ap, ap_d, ap_t, pe, pe_d, pe_t, new_ref, nrp, nvp, enter = r
            pno_v.append(v)
            pno_curve.append(np.array(curve, dtype=float))
            pno_sc.append([ap[0], ap[1], ap_d, ap_t, pe[0], pe[1], pe_d, pe_t, float(new_ref),
                           nrp[0], nrp[1], nvp[0], nvp[1], float(enter)])
        out.update({"s%d_pno_v" % k: np.array(pno_v, dtype=np.int64),
                    "s%d_pno_curve" % k: np.array(pno_curve, dtype=float).reshape(len(pno_v), P, 2),
                    "s%d_pno_sc" % k: np.array(pno_sc, dtype=float).reshape(len(pno_v), 14)})
        # rotate(): active vessel 3 accelerates at warp 1, everything stops at higher warp
        r_in = (np.array(pv.rot_angle, dtype=float), np.array(pv.rot_speed, dtype=float))
        ang1 = np.array(pv.rotate(1, 3, 1.0), dtype=float)
        ang2 = np.array(pv.rotate(1, 3, -1.0), dtype=float)
        sp2 = np.array(pv.rot_speed, dtype=float)
        ang3 = np.array(pv.rotate(5, 3, 1.0), dtype=float)
        sp3 = np.array(pv.rot_speed, dtype=float)
        out.update({"s%d_rot_angle_in" % k: r_in[0], "s%d_rot_speed_in" % k: r_in[1],
                    "s%d_rot_angle1" % k: ang1, "s%d_rot_angle2" % k: ang2, "s%d_rot_speed2" % k: sp2,
                    "s%d_rot_angle3" % k: ang3, "s%d_rot_speed3" % k: sp3})
        # keep some spin for the next snapshot
        pv.rot_speed = np.array(pv.rot_speed, dtype=float) + rng.uniform(-0.05, 0.05, nv)
        print("  vessel_draw snapshot %d (tick %d): types %s, %d selected, %d predict_next_orbit" % (
            k, tick, np.bincount(itype, minlength=4).tolist(), int(np.sum(~np.isnan(sel[:, 0]))),
            len(pno_v)))
        snaps += 1
    out["n_snap"] = snaps
    save("vessel_draw", **out)
    print("  vessel_draw coverage (type, ell, ccw): %s" % sorted(cover.items()))


def gen_mapfiles(ns, conf, bd, bod):
    """N4 (savefile I/O): maps written by the reference's peripherals.save_file from seeded arrays
    (Newtonian, and Keplerian with vessels) and what its load_file returns for them."""
    rng = np.random.default_rng(20240625)
    os.chdir(rh._WORKDIR)                       # save_file creates Maps/ and Saves/ in the CWD
    mdir = os.path.join(OUT, "maps")
    os.makedirs(mdir, exist_ok=True)
    out = {}

    def bodies(n):
        atm = rng.uniform(0, 1, n) < 0.4
        names = np.array(["Body %d" % i for i in range(n)])
        names[3] = names[2]                     # duplicate name -> "Body 2 1" on save
        return {"name": names, "mass": 10 ** rng.uniform(-1, 4, n), "den": rng.uniform(500, 6000, n),
                "color": rng.integers(0, 256, (n, 3)), "atm_pres0": np.where(atm, rng.uniform(0.1, 90, n), 0.0),
                "atm_scale_h": np.where(atm, rng.uniform(0.001, 0.1, n), 0.0),
                "atm_den0": np.where(atm, rng.uniform(0.1, 70, n), 0.0)}

    def dump(tag, res):
        gd, cf, b, bo, v, vo = res
        out[tag + "_game"] = np.array([gd["name"], gd["date"], str(gd["time"]), str(gd["vessel"])])
        out[tag + "_conf_keys"] = np.array(list(cf.keys()))
        out[tag + "_conf_vals"] = np.array([float(x) for x in cf.values()])
        for k, a in b.items():
            out[tag + "_b_" + k] = np.asarray(a)
        for k, a in bo.items():
            out[tag + "_bo_" + k] = np.asarray(a)
        for k, a in v.items():
            out[tag + "_v_" + k] = np.asarray(a)
        for k, a in vo.items():
            out[tag + "_vo_" + k] = np.asarray(a)

    cfg = dict(conf)
    # Newtonian map
    n = 40
    b = bodies(n)
    bo = {"pos": rng.uniform(-1e5, 1e5, (n, 2)), "vel": rng.uniform(-3, 3, (n, 2))}
    gd = {"name": "Newtonian test", "date": "01.01.2024 12:00", "time": 1234.5, "vessel": None}
    path = os.path.join(mdir, "newton.ini")
    ns.peripherals.save_file(path, gd, cfg, b, bo)
    dump("newton", ns.peripherals.load_file(path))
    # Keplerian savegame with vessels
    n, nv = 30, 12
    b = bodies(n)
    bo = {"a": rng.uniform(1e3, 1e5, n), "ecc": rng.uniform(0, 0.9, n), "pe_arg": rng.uniform(0, 6.28, n),
          "ma": rng.uniform(0, 6.28, n), "ref": rng.integers(0, n, n), "dir": rng.choice([-1.0, 1.0], n)}
    v = {"name": np.array(["Vessel %d" % i for i in range(nv)]), "mass": rng.uniform(1, 50, nv),
         "rot_angle": rng.uniform(0, 6.28, nv), "rot_acc": rng.uniform(0, 0.1, nv),
         "sprite": np.array(["sprite%d.png" % (i % 3) for i in range(nv)])}
    vecc = np.where(rng.uniform(0, 1, nv) < 0.4, rng.uniform(1.1, 3, nv), rng.uniform(0, 0.9, nv))
    vo = {"a": rng.uniform(100, 5000, nv) * np.where(vecc > 1, -1, 1), "ecc": vecc,
          "pe_arg": rng.uniform(0, 6.28, nv), "ma": rng.uniform(-3, 6, nv), "ref": rng.integers(0, n, nv),
          "dir": rng.choice([-1.0, 1.0], nv)}
    gd = {"name": "Kepler test", "date": "02.02.2024 08:30", "time": 99.0, "vessel": 3}
    path = os.path.join(mdir, "kepler.ini")
    ns.peripherals.save_file(path, gd, cfg, b, bo, v, vo)
    dump("kepler", ns.peripherals.load_file(path))
    # the built-in map: arrays only (the file itself stays in the reference tree)
    dump("solar", ns.peripherals.load_file(ns.solar_system_ini))
    save("mapfiles", **out)


def gen_convert_vel(ns):
    """convert.kepler_to_velocity / velocity_to_kepler (all failsafe modes) and phys_shared.orb2xy
    of the live reference on seeded inputs."""
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    rng = np.random.default_rng(20240627)
    n = 300
    hyp = rng.uniform(0, 1, n) < 0.4
    ecc = np.where(hyp, rng.uniform(1.05, 3.0, n), rng.uniform(0.01, 0.9, n))
    a = rng.uniform(500, 5000, n) * np.where(hyp, -1, 1)
    pea = rng.uniform(0, 2 * np.pi, n)
    u = rng.uniform(1e3, 1e5, n)
    dr = rng.choice([-1.0, 1.0], n)
    ea = np.where(hyp, rng.uniform(-2, 2, n), rng.uniform(0, 2 * np.pi, n))
    ea[:3] = 0.0                                   # orb2xy: (0, 0) without the reference position
    f = a * ecc
    b = np.sqrt(np.abs(f**2 - a**2))
    ref_pos = rng.uniform(-1e4, 1e4, (n, 2))
    xy = np.array([ns.phys_shared.orb2xy(a[i], b[i], f[i], ecc[i], pea[i], ref_pos[i], ea[i])
                   for i in range(n)])
    rel = xy - ref_pos
    rel[:3] = rng.uniform(500, 900, (3, 2))
    with np.errstate(all="ignore"):
        vel = np.array([convert.kepler_to_velocity(rel[i], a[i], ecc[i], pea[i], u[i], dr[i])
                        for i in range(n)])
    out = dict(a=a, b=b, f=f, ecc=ecc, pea=pea, u=u, dr=dr, ea=ea, ref_pos=ref_pos, xy=xy, rel=rel,
               vel=vel)
    # velocity_to_kepler on random state vectors (bound and unbound), three failsafe modes
    m = 400
    pos = rng.uniform(-5e3, 5e3, (m, 2))
    uu = rng.uniform(1e3, 1e5, m)
    speed = np.sqrt(uu / np.linalg.norm(pos, axis=1)) * rng.uniform(0.3, 2.2, m)
    ang = rng.uniform(0, 2 * np.pi, m)
    vv = speed[:, None] * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    for fs in (0, 1, 2):
        res = np.array([convert.velocity_to_kepler(pos[i].copy(), vv[i].copy(), uu[i], failsafe=fs)
                        for i in range(m)], dtype=float)
        out["v2k_fs%d" % fs] = res
    out.update(v2k_pos=pos, v2k_vel=vv, v2k_u=uu)
    save("convert_vel", **out)


def bits_xor(a):
    """Order-independent fingerprint of an array's exact bits (tests rebuild the inputs the
    live reference saw and must prove they are the same bits)."""
    return np.bitwise_xor.reduce(np.ascontiguousarray(a, dtype=np.float64).reshape(-1).view(np.uint64))


def gen_fullsize(ns):
    """Round-2 fixtures at the sizes SURVEY 8d names ("next" item 1):
      * collision_k64: config 5 teacher-forced at N = 4096 over K = 64 ticks -- the genuine
        reference's check_collision (phys_editor.py:547-551) on a seeded kinematic collapse
        (synth.kinematic_collapse: pos += vel, merges by the documented rule), (del, add) per tick;
      * editor_curve_conics: C7 (phys_editor.py:519-544) on ellipses, hyperbolas AND exact
        ecc == 1 parabolas (:532-534);
      * body_move_deep: C1 (phys_body.py:323-346) on a 131 072-body hierarchy of depth 3 with
        children that come before their parent in argsort(coi)[::-1]; 4096 sampled rows kept."""
    sys.path.insert(0, os.path.dirname(HERE))
    from volatilespace_b200 import synth
    from volatilespace_b200.phys_editor import body_radius
    # ---- collision_k64
    s = synth.kinematic_collapse()
    ph = ns.phys_editor.Physics()

    def ref_check(mass, pos, rad):
        ph.mass, ph.pos, ph.rad = mass, pos, rad
        return ph.check_collision()
    pairs, fp = [], []
    for t, mass, pos, rad, pair in synth.kinematic_collapse_run(s, 64, ref_check, body_radius):
        pairs.append([-1, -1] if pair[0] is None else [int(pair[0]), int(pair[1])])
        fp.append([len(mass), int(bits_xor(pos)), int(bits_xor(rad)), int(bits_xor(mass))])
    save("collision_k64", pairs=np.array(pairs, dtype=np.int64), fingerprint=np.array(fp, dtype=np.uint64),
         n=np.int64(len(s["mass"])), ticks=np.int64(64))
    # ---- editor_curve_conics
    rng = np.random.default_rng(20240607)
    n, P = 96, 48
    kind = np.arange(n) % 3                     # 0 ellipse, 1 hyperbola, 2 parabola
    ang = rng.uniform(0, 2 * np.pi, n)
    emag = np.where(kind == 0, rng.uniform(0.0, 0.95, n), rng.uniform(1.05, 3.0, n))
    ecc_v = np.stack([emag * np.cos(ang), emag * np.sin(ang)], axis=1)
    unit = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0], [0.6, 0.8], [-0.8, 0.6]])
    par = np.flatnonzero(kind == 2)
    ecc_v[par] = unit[np.arange(len(par)) % len(unit)]
    assert all(ns.phys_shared.mag(ecc_v[i]) == 1.0 for i in par)
    a = rng.uniform(500, 5e4, n) * np.where(kind == 1, -1.0, 1.0)
    f = a * np.where(kind == 2, 1.0, emag)
    b = np.sqrt(np.abs(f * f - a * a))
    pe = rng.uniform(0, 2 * np.pi, n)
    parents = rng.integers(0, n, n)
    parents[0] = 0
    pos = rng.uniform(-1e5, 1e5, (n, 2))
    ph = ns.phys_editor.Physics()
    rh.set_curve_points(ns, ph, P)
    ph.ecc_v, ph.semi_major, ph.semi_minor, ph.focus, ph.periapsis_arg = ecc_v, a, b, f, pe
    ph.parents, ph.pos = parents, pos
    save("editor_curve_conics", P=np.int64(P), ecc_v=ecc_v, semi_major=a, semi_minor=b, focus=f,
         periapsis_arg=pe, parents=parents, pos=pos, curves=ph.curve())
    # ---- body_move_deep
    h = synth.kepler_hierarchy()
    n = len(h["a"])
    pb = ns.phys_body.Physics()
    for k in ("a", "b", "f", "ecc", "pea", "n", "coi", "ref"):
        setattr(pb, k, h[k].copy())
    pb.dr, pb.ma, pb.ea, pb.pos = h["dr"].copy(), h["ma"].copy(), h["ea"].copy(), h["pos"].copy()
    sample = np.sort(np.random.default_rng(3).choice(n, 4096, replace=False))
    out = dict(n=np.int64(n), sample=sample,
               inputs_fp=np.array([int(bits_xor(h[k])) for k in ("a", "b", "f", "ecc", "pea", "n", "coi",
                                                                 "dr", "ma", "ea")], dtype=np.uint64),
               ref_fp=np.int64(np.bitwise_xor.reduce(h["ref"])))
    for step, warp in enumerate((1, 3, 2)):
        pos, ma, ea = pb.move(warp)
        out.update({"pos%d" % step: pos[sample].copy(), "ma%d" % step: ma[sample].copy(),
                    "ea%d" % step: ea[sample].copy()})
    save("body_move_deep", **out)


def editor_hierarchy(n=4400, seed=20240612, coi_coef=0.7):
    """A MODE_REF system above the 4096-body switch of the production kernels (the O(N) candidate
    passes for find_parents / check_collision and the fixed-point rounds of simplified_orbit_coi):
    root 0, 60 planets on exactly circular orbits, ~1500 moons inside the planets' COIs, 120
    moonlets around heavy moons (depth 3), 400 bodies within 1.2 % of a planet's COI edge moving
    radially (parent changes in both directions within a few frames), free bodies, and four planted
    overlaps (free-free twice, moon-moon, planet-moon) that merge on the first frames."""
    rng = np.random.default_rng(seed)
    M0 = 1.0e7
    mass = rng.uniform(0.5, 2.0, n)
    den = np.full(n, 3.0e5)
    pos = np.zeros((n, 2))
    vel = np.zeros((n, 2))
    mass[0] = M0

    def circ(center, cvel, m_c, d, th, speed_mult=1.0):
        u = np.stack([np.cos(th), np.sin(th)], axis=-1)
        t = np.stack([-np.sin(th), np.cos(th)], axis=-1)
        sp = np.sqrt(m_c / d) * speed_mult
        return center + d[..., None] * u, cvel + sp[..., None] * t, u
    npl = 60
    pl = np.arange(1, 1 + npl)
    mass[pl] = rng.uniform(2.0e3, 5.0e4, npl)
    r_pl = np.sort(rng.uniform(2.0e5, 3.0e6, npl))
    pos[pl], vel[pl], _ = circ(np.zeros(2), np.zeros(2), M0, r_pl, rng.uniform(0, 2 * np.pi, npl))
    coi_pl = r_pl * (mass[pl] / M0) ** coi_coef
    k = 1 + npl
    # moons
    nmoon = 1500
    mo = np.arange(k, k + nmoon)
    par = rng.integers(0, npl, nmoon)
    d = coi_pl[par] * rng.uniform(0.05, 0.6, nmoon)
    pos[mo], vel[mo], _ = circ(pos[pl][par], vel[pl][par], mass[pl][par], d, rng.uniform(0, 2 * np.pi, nmoon),
                               rng.normal(1.0, 0.05, nmoon))
    heavy = mo[:120]
    mass[heavy] = rng.uniform(100.0, 400.0, 120)
    k += nmoon
    # moonlets around the heavy moons (their COI: ~d * (m / m_planet)^0.7)
    ml = np.arange(k, k + 120)
    d_h = np.linalg.norm(pos[heavy] - pos[pl][par[:120]], axis=1)
    coi_h = d_h * (mass[heavy] / mass[pl][par[:120]]) ** coi_coef
    mass[ml] = rng.uniform(0.01, 0.05, 120)
    den[ml] = 3.0e7
    pos[ml], vel[ml], _ = circ(pos[heavy], vel[heavy], mass[heavy], coi_h * rng.uniform(0.15, 0.5, 120),
                               rng.uniform(0, 2 * np.pi, 120))
    k += 120
    # COI-edge crossers
    ncr = 400
    cr = np.arange(k, k + ncr)
    cp = rng.integers(0, npl, ncr)
    inside = rng.uniform(size=ncr) < 0.5
    d = coi_pl[cp] * (1.0 + np.where(inside, -1.0, 1.0) * rng.uniform(0.0, 0.012, ncr))
    pos[cr], vel[cr], u = circ(pos[pl][cp], vel[pl][cp], mass[pl][cp], d, rng.uniform(0, 2 * np.pi, ncr), 0.7)
    vel[cr] += u * (np.where(inside, 1.0, -1.0) * coi_pl[cp] * rng.uniform(0.0008, 0.002, ncr))[:, None]
    mass[cr] = rng.uniform(0.01, 0.5, ncr)
    k += ncr
    # free bodies around the root
    fr = np.arange(k, n)
    pos[fr], vel[fr], _ = circ(np.zeros(2), np.zeros(2), M0, rng.uniform(1.0e5, 3.5e6, len(fr)),
                               rng.uniform(0, 2 * np.pi, len(fr)), rng.normal(1.0, 0.05, len(fr)))
    # planted overlaps: radius = 179 cbrt(3 m / (4 pi den)) ~ 1.7 for m = 1
    def touch(i, j, dist):
        pos[j] = pos[i] + np.array([dist, 0.0])
        vel[j] = vel[i]
    touch(fr[10], fr[900], 1.0)
    touch(fr[400], fr[77], 0.5)
    touch(mo[700], mo[1300], 0.8)
    touch(pl[17], mo[300], 20.0)          # planet radius ~ 60
    return mass, den, np.ascontiguousarray(pos), np.ascontiguousarray(vel)


def gen_editor_large(ns):
    """MODE_REF frames of the GENUINE reference at 4400 bodies -- the size class in which the
    library runs its O(N) candidate passes and fixed-point rounds instead of the small-system
    kernels (phys_editor.py:325-378, 547-565, 844-857, 250-290).  Parents of every body on every
    frame, deleted indices, body counts; the float state on sampled rows after load, after the
    merges and at the end."""
    import time
    mass, den, pos, vel = editor_hierarchy()
    n = len(mass)
    t0 = time.time()
    ph = rh.editor_from_arrays(ns, mass, den, pos, vel)
    print("editor_large: load %.1f s" % (time.time() - t0))
    out = dict(mass=mass, den=den, pos0=pos, vel0=vel, frames=np.int64(8))
    rng = np.random.default_rng(5)

    def sampled(tag):
        m = len(ph.mass)
        rows = np.sort(rng.choice(m, 1200, replace=False))
        out[tag + "rows"] = rows
        for k2, v in snap(ph, tag).items():
            out[k2] = v if k2.endswith("parents") else v[rows]
    sampled("f0_")
    dels, counts = [], []
    for fr in range(1, 9):
        t0 = time.time()
        d = rh.editor_frame(ph)
        dels.append(d[0] if d else -1)
        counts.append(len(ph.mass))
        out["parents_f%d" % fr] = np.array(ph.parents, dtype=np.int16)
        print("editor_large: frame %d %.1f s, deleted %s, %d bodies" % (fr, time.time() - t0, d, len(ph.mass)))
        if fr in (4, 8):
            sampled("f%d_" % fr)
    out["deleted"] = np.array(dels)
    out["counts"] = np.array(counts)
    changes = sum(int((out["parents_f%d" % (f + 1)] != out["parents_f%d" % f]).sum())
                  for f in range(5, 8))
    print("editor_large: merges %d, parent changes on frames 6-8: %d, depth-2 parents %d" % (
        (np.array(dels) >= 0).sum(), changes, int((out["f0_parents"] > 60).sum())))
    assert n - counts[-1] >= 3 and changes >= 20 and (out["f0_parents"] > 60).sum() >= 60
    save("editor_large", **out)


def gen_convert_large(ns, conf):
    """N1 at 4400 bodies (above the 4096-body switch: the fixed-point rounds over the candidate
    pass instead of the blocked sweep): the GENUINE convert.to_kepler (convert.py:132-261) on the
    editor_large system with every index but the root scrambled.  refs of all bodies, elements on
    sampled rows; the inputs are editor_large's arrays + the stored permutation."""
    import time
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from volatilespace.physics import convert
    sys.path.remove(rh.REFERENCE_ROOT)
    mass, den, pos, vel = editor_hierarchy()
    n = len(mass)
    # the four planted partners share their partner's velocity exactly; next to a planet that is a
    # radial orbit on which the reference divides by zero (convert.py:252): give them a push
    _, first, counts = np.unique(vel, axis=0, return_index=True, return_counts=True)
    dup = np.setdiff1d(np.arange(n), first)
    assert len(dup) == 4
    vel = vel.copy()
    vel[dup] += np.array([0.05, 0.3])
    rng = np.random.default_rng(20240613)
    perm = np.concatenate(([0], 1 + rng.permutation(n - 1)))
    t0 = time.time()
    k = convert.to_kepler(mass[perm], {"pos": pos[perm].copy(), "vel": vel[perm].copy()}, conf["gc"],
                          conf["coi_coef"])
    print("convert_large: to_kepler %.1f s, %d refs != 0" % (time.time() - t0, int((np.array(k["ref"]) != 0).sum())))
    rows = np.sort(rng.choice(n, 1200, replace=False))
    out = dict(perm=perm.astype(np.int16), rows=rows, pushed=dup.astype(np.int16), push=np.array([0.05, 0.3]), k_ref=np.array(k["ref"], dtype=np.int16),
               gc=conf["gc"], coi_coef=conf["coi_coef"])
    for key in ("a", "ecc", "pe_arg", "ma", "dir"):
        out["k_" + key] = np.array(k[key])[rows]
    assert (out["k_ref"] != 0).sum() > 300
    save("convert_large", **out)


def main():
    """All fixtures, or only the ones named on the command line (e.g. `vessel_points`)."""
    ns = rh.load()
    only = set(sys.argv[1:])
    want = lambda name: not only or name in only
    conf, bd, bod = gen_solar(ns) if want("solar") else gen_solar_inputs(ns)
    if want("editor_clouds"): gen_editor_clouds(ns)
    if want("pair_kernels"): gen_pair_kernels(ns)
    if want("allpairs"): gen_allpairs(ns)
    if want("kepler_solvers"): gen_kepler_solvers(ns)
    if want("game"): gen_game(ns, conf, bd, bod)
    if want("culling"): gen_culling(ns, conf, bd, bod)
    if want("body_data"): gen_body_data(ns, conf, bd, bod)
    if want("editor_edits"): gen_editor_edits(ns)
    if want("convert"): gen_convert(ns, conf, bd, bod)
    if want("intersect"): gen_intersect(ns)
    if want("vessel_points"): gen_vessel_points(ns, conf, bd, bod)
    if want("vessel_events"): gen_vessel_events(ns, conf, bd, bod)
    if want("vessel_draw"): gen_vessel_draw(ns, conf, bd, bod)
    if want("mapfiles"): gen_mapfiles(ns, conf, bd, bod)
    if want("convert_vel"): gen_convert_vel(ns)
    if want("fullsize"): gen_fullsize(ns)
    if want("editor_large"): gen_editor_large(ns)
    if want("convert_large"): gen_convert_large(ns, conf)


if __name__ == "__main__":
    main()
