"""Seeded synthetic inputs for the BASELINE.json configs.

Pure numpy, host side; used by bench.py and the tests.  The reference is 2-D
(SURVEY F2), so "Plummer sphere" / "rotating disk" are 2-D distributions.
"""
import numpy as np


def plummer2d(n=65536, seed=20240601, a=1.0e5, gc=1.0):
    """Config 2: projected Plummer disk, surface density ~ (1+R^2/a^2)^-2."""
    rng = np.random.default_rng(seed)
    mass = rng.uniform(0.5, 2.0, n)
    u = rng.uniform(0.0, 0.99, n)
    R = a * np.sqrt(u / (1.0 - u))
    th = rng.uniform(0.0, 2 * np.pi, n)
    heavy = int(np.argmax(mass))
    mass[[0, heavy]] = mass[[heavy, 0]]          # body 0 = heaviest, at the origin
    R[0] = 0.0
    pos = np.stack([R * np.cos(th), R * np.sin(th)], axis=1)
    mtot = mass.sum()
    menc = mtot * R * R / (R * R + a * a)
    with np.errstate(divide="ignore", invalid="ignore"):
        sp = np.where(R > 0, np.sqrt(gc * menc / np.maximum(R, 1e-300)), 0.0)
    sp *= rng.normal(1.0, 0.1, n)
    vel = np.stack([-sp * np.sin(th), sp * np.cos(th)], axis=1)
    vel[0] = 0.0
    den = np.full(n, 3000.0)
    return dict(mass=mass, den=den, pos=np.ascontiguousarray(pos), vel=np.ascontiguousarray(vel),
                gc=gc, rad_mult=179.0)


def rotating_disk(n=1048576, seed=20240602, gc=1.0):
    """Config 3: central mass 1e9 + light bodies on circular counter-clockwise orbits."""
    rng = np.random.default_rng(seed)
    mass = rng.uniform(0.5, 2.0, n)
    mass[0] = 1.0e9
    r = rng.uniform(1.0e5, 1.0e7, n)
    th = rng.uniform(0.0, 2 * np.pi, n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    sp = np.sqrt(gc * mass[0] / r)
    vel = np.stack([-sp * np.sin(th), sp * np.cos(th)], axis=1)
    pos[0] = 0.0
    vel[0] = 0.0
    den = np.full(n, 3000.0)
    return dict(mass=mass, den=den, pos=np.ascontiguousarray(pos), vel=np.ascontiguousarray(vel),
                gc=gc, rad_mult=179.0)


def collapsing_cluster(n=262144, seed=20240604, gc=1.0, radius=2.0e5):
    """Config 5: cold uniform disk + one 1e6 central body; radii overlap as it collapses."""
    rng = np.random.default_rng(seed)
    mass = rng.uniform(0.5, 2.0, n)
    mass[0] = 1.0e6
    r = radius * np.sqrt(rng.uniform(0.0, 1.0, n))
    th = rng.uniform(0.0, 2 * np.pi, n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    pos[0] = 0.0
    vel = np.zeros((n, 2))
    den = np.full(n, 3000.0)
    return dict(mass=mass, den=den, pos=np.ascontiguousarray(pos), vel=vel, gc=gc, rad_mult=179.0)


def calc_orb(kind, ref, ref_mass, body_mass, gc, coi_coef, a, ecc):
    """C8, vectorised: calc_orb_one of phys_body.py:27-60 (kind='body', index 0 is the
    root and gets zeros) or phys_vessel.py:82-110 (kind='vessel').  Host numpy, same
    expressions as the reference's scalar code.  Returns dict b,f,coi,pe_d,ap_d,period,n,u."""
    a = np.asarray(a, dtype=np.float64)
    ecc = np.asarray(ecc, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.int64)
    ref_mass = np.asarray(ref_mass, dtype=np.float64)
    u = gc * ref_mass[ref]
    f = a * ecc
    b = np.sqrt(np.abs(f**2 - a**2))
    with np.errstate(divide="ignore", invalid="ignore"):
        period_ell = 2 * np.pi * np.sqrt(a**3 / u)
        n_ell = 2 * np.pi / period_ell
        n_hyp = np.sqrt(u / np.abs(a)**3)
        period_circ = period_ell / 10
        n_circ = 2 * np.pi / period_circ
    ell = (ecc != 0) & (ecc < 1)
    circ = ecc == 0
    pe_d = np.where((ecc == 1), a, a * (1 - ecc))
    ap_d = np.where(ell, a * (1 + ecc), 0.0)
    period = np.where(ell, period_ell, np.where(circ, period_circ, 0.0))
    n = np.where(ell, n_ell, np.where(circ, n_circ, n_hyp))
    out = dict(b=b, f=f, pe_d=pe_d, ap_d=ap_d, period=period, n=n, u=u)
    if kind == "body":
        body_mass = np.asarray(body_mass, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            coi = np.where(a > 0, a * (body_mass / ref_mass[ref])**coi_coef, 0.0)
        out["coi"] = coi
        for k in out:                       # root: all zeros (phys_body.py:60)
            out[k] = np.array(out[k], dtype=np.float64)
            out[k][0] = 0.0
    else:
        out["coi"] = np.zeros(len(a))
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in out.items()}


def kepler_objects(n=4194304, seed=20240603, gc=1.0, central_mass=1.0e6, quirk=0):
    """Config 4: n objects on mixed elliptic (75 %) / hyperbolic (25 %) conics around one
    body (ref = 0), vessel semantics.  `quirk` objects are forced into the ENRKE
    quirk region (ecc > 0.99, folded mean anomaly < 0.0045)."""
    rng = np.random.default_rng(seed)
    hyp = rng.uniform(0, 1, n) < 0.25
    ecc = np.where(hyp, rng.uniform(1.05, 3.0, n), rng.uniform(0.0, 0.95, n))
    a = rng.uniform(5.0e3, 5.0e5, n) * np.where(hyp, -1.0, 1.0)
    pea = rng.uniform(0.0, 2 * np.pi, n)
    ma = np.where(hyp, rng.uniform(-3.0, 3.0, n), rng.uniform(0.0, 2 * np.pi, n))
    dr = rng.choice([-1.0, 1.0], n)
    if quirk:
        q = slice(0, quirk)
        ecc[q] = rng.uniform(0.9901, 0.9999, quirk)
        a[q] = np.abs(a[q])
        m = rng.uniform(1e-5, 0.0044, quirk)
        ma[q] = np.where(rng.uniform(0, 1, quirk) < 0.5, m, 2 * np.pi - m)
    ref = np.zeros(n, dtype=np.int64)
    body_mass = np.array([central_mass])
    body_pos = np.zeros((1, 2))
    orb = calc_orb("vessel", ref, body_mass, None, gc, 0.7, a, ecc)
    ea = np.where(ecc >= 1, ma, 0.0)           # hyperbolic warm start (phys_vessel.py:221)
    return dict(a=a, ecc=ecc, pea=pea, ma=ma, dr=dr, ref=ref, b=orb["b"], f=orb["f"],
                n=orb["n"], ea=ea, body_pos=body_pos, body_mass=body_mass, gc=gc)


def kepler_bodies(n=4194304, seed=20240603, gc=1.0, central_mass=1.0e6, quirk=0, coi_coef=0.7):
    """Config 4, body-style variant (C1, phys_body.py:323-346): the same conics as
    kepler_objects, as BODIES of one root (index 0, all-zero elements, phys_body.py:60) with
    ref = 0 everywhere, light masses and the `coi` that orders the reference's loop."""
    k = kepler_objects(n, seed=seed, gc=gc, central_mass=central_mass, quirk=quirk)
    rng = np.random.default_rng(seed + 1)
    mass = rng.uniform(0.5, 2.0, n)
    mass[0] = central_mass
    a, ecc = k["a"].copy(), k["ecc"].copy()
    a[0] = ecc[0] = 0.0
    orb = calc_orb("body", k["ref"], mass, mass, gc, coi_coef, a, ecc)
    ma, pea, dr = k["ma"].copy(), k["pea"].copy(), k["dr"].copy()
    ma[0] = pea[0] = 0.0
    ea = np.where(ecc >= 1, ma, 0.0)
    return dict(a=a, ecc=ecc, pea=pea, ma=ma, dr=dr, ref=k["ref"], b=orb["b"], f=orb["f"],
                n=orb["n"], coi=orb["coi"], ea=ea, mass=mass, pos=np.zeros((n, 2)), gc=gc)


def kepler_hierarchy(n=131072, seed=20240605, gc=1.0, coi_coef=0.7):
    """C1 on a deep hierarchy: root 0 -> planets -> moons -> moonlets (depth 3), indices
    shuffled so that `ref` points both ways in index order, 8 % hyperbolic bodies on every
    level.  A hyperbolic body has coi = 0 (calc_orb_one, phys_body.py:47-50), so its children
    come BEFORE it in argsort(coi)[::-1] and are composed with its position of the previous
    tick -- the loop-order behaviour of phys_body.py:328-345 that the device path must keep."""
    rng = np.random.default_rng(seed)
    n1 = max(2, n // 128)
    n2 = max(2, n // 12)
    n3 = n - 1 - n1 - n2
    perm = np.concatenate(([0], 1 + rng.permutation(n - 1)))       # new index of logical body k
    lvl = np.concatenate(([0], np.full(n1, 1), np.full(n2, 2), np.full(n3, 3)))
    lref = np.zeros(n, dtype=np.int64)                             # logical parent
    lref[1 + n1:1 + n1 + n2] = rng.integers(1, 1 + n1, n2)
    lref[1 + n1 + n2:] = rng.integers(1 + n1, 1 + n1 + n2, n3)
    scale_a = np.array([0.0, 1.0e8, 1.0e4, 30.0])[lvl]
    scale_m = np.array([1.0e9, 1.0e4, 10.0, 0.05])[lvl]
    hyp = rng.uniform(0, 1, n) < 0.08
    hyp[0] = False
    ecc_l = np.where(hyp, rng.uniform(1.05, 2.5, n), rng.uniform(0.0, 0.9, n))
    a_l = scale_a * rng.uniform(0.1, 1.0, n) * np.where(hyp, -1.0, 1.0)
    mass_l = scale_m * rng.uniform(0.5, 2.0, n)
    pea_l = rng.uniform(0, 2 * np.pi, n)
    ma_l = np.where(hyp, rng.uniform(-2.0, 2.0, n), rng.uniform(0, 2 * np.pi, n))
    dr_l = rng.choice([-1.0, 1.0], n)
    a_l[0] = ecc_l[0] = pea_l[0] = ma_l[0] = 0.0
    out = {}
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)                                       # logical body at new index
    for name, arr in (("a", a_l), ("ecc", ecc_l), ("mass", mass_l), ("pea", pea_l), ("ma", ma_l),
                      ("dr", dr_l), ("level", lvl)):
        out[name] = np.ascontiguousarray(arr[inv])
    out["ref"] = np.ascontiguousarray(perm[lref[inv]])
    orb = calc_orb("body", out["ref"], out["mass"], out["mass"], gc, coi_coef, out["a"], out["ecc"])
    out.update(b=orb["b"], f=orb["f"], n=orb["n"], coi=orb["coi"],
               ea=np.where(out["ecc"] >= 1, out["ma"], 0.0), pos=np.zeros((n, 2)), gc=gc)
    return out


def kinematic_collapse(n=4096, seed=20240606, radius=4.0e4, rad_mult=179.0):
    """Teacher-forced collision input (config 5 at the size the genuine reference can run): a
    uniform disk whose bodies fall radially inward at constant speed, advanced by pos += vel only
    (elementwise IEEE adds: every machine regenerates the same bits), so that the states the
    live reference saw while the fixture was written can be rebuilt inside the test."""
    rng = np.random.default_rng(seed)
    mass = rng.uniform(0.5, 2.0, n)
    mass[0] = 200.0
    r = radius * np.sqrt(rng.uniform(0.0, 1.0, n))
    th = rng.uniform(0.0, 2 * np.pi, n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    pos[0] = 0.0
    vel = -pos * rng.uniform(0.002, 0.01, (n, 1))
    den = np.full(n, 3000.0)
    return dict(mass=mass, den=den, pos=np.ascontiguousarray(pos), vel=np.ascontiguousarray(vel),
                rad_mult=rad_mult)


def kinematic_collapse_run(s, ticks, check_collision, body_radius):
    """Drives kinematic_collapse for `ticks` ticks: pos += vel; rad = body_radius(...);
    (del, add) = check_collision(mass, pos, rad); documented inelastic merge
    (documentation/orbital-physics/09-collisions.txt:3-6: mass and momentum add up, the survivor
    keeps its position) and np.delete of the deleted row.  Yields (tick, mass, pos, rad, pair)
    BEFORE the merge is applied; `pair` is what check_collision returned."""
    mass, den, pos, vel = (s[k].copy() for k in ("mass", "den", "pos", "vel"))
    for t in range(ticks):
        pos += vel
        rad = body_radius(mass, den, s["rad_mult"])
        pair = check_collision(mass, pos, rad)
        yield t, mass, pos, rad, pair
        if pair[0] is not None:
            d, a = int(pair[0]), int(pair[1])
            mr = mass[d] + mass[a]
            vel[a] = (vel[a] * mass[a] + vel[d] * mass[d]) / mr
            mass[a] = mr
            mass, den = np.delete(mass, d), np.delete(den, d)
            pos, vel = np.delete(pos, d, axis=0), np.delete(vel, d, axis=0)
