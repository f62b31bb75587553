"""GPU parity tests for part C (game-mode Keplerian propagation + orbit curves).

Tolerances: ea, ma, pos rel 1e-12 (iteration counts may
differ by one near the solver tolerance -> 1e-10 on ea where stated), curve
points rel 1e-12 of the curve's scale.
"""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from volatilespace_b200.device import Device
    d = Device(0)
    yield d
    d.close()


def test_solve_kepler_ell_golden_and_known_answers(dev):
    g = golden("kepler_solvers")
    got = dev.host_solve_kepler_ell(g["ecc"], g["ma"], 1e-10)
    # the Newton exit test may fire one iteration apart: both sides are within tol of the root
    np.testing.assert_allclose(got, g["ell"], rtol=1e-9, atol=1e-10)
    close = np.isclose(got, g["ell"], rtol=1e-13, atol=1e-14)
    assert close.mean() > 0.97
    got6 = dev.host_solve_kepler_ell(g["ecc"], g["ma"], 1e-6)
    np.testing.assert_allclose(got6, g["ell_tol6"], rtol=1e-5, atol=1e-6)
    known = {(0.1, 1.0): 1.0885977523978936, (0.5, 3.0): 3.0471507747023945,
             (0.9, 0.1): 0.6308435275631534, (0.0167, 4.5): 4.4837346626129175,
             (0.7, 6.0): 5.512220917883771, (0.995, 0.001): 0.07835}
    e = np.array([k[0] for k in known])
    m = np.array([k[1] for k in known])
    np.testing.assert_allclose(dev.host_solve_kepler_ell(e, m, 1e-10), list(known.values()),
                               rtol=1e-12)


def test_enrke_quirk_branch_exact(dev, oracle):
    """ecc > 0.99 and folded ma < 0.0045: the reference's non-converging bisection (F8b)."""
    rng = np.random.default_rng(1)
    ecc = rng.uniform(0.9901, 0.99999, 4000)
    m = rng.uniform(0.0, 0.0045, 4000)
    ma = np.where(rng.uniform(0, 1, 4000) < 0.5, m, 2 * np.pi - m)
    got = dev.host_solve_kepler_ell(ecc, ma, 1e-10)
    want = oracle.solve_kepler_ell(ecc, ma, 1e-10)
    np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-16)


def test_newton_hyp_golden(dev):
    g = golden("kepler_solvers")
    got = dev.host_newton_root_kepler_hyp(g["hecc"], g["hma"], g["hguess"])
    want = g["hyp"]
    ok = np.isfinite(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(
        dev.host_newton_root_kepler_hyp([1.5, 3.0, 1.01], [2.0, -4.0, 0.5], [2.0, 0.0, 0.5]),
        [-1.6126858097584946, 1.3408203817380824, -1.379754746227187], rtol=1e-12)


def test_curve_points_and_move_to_golden(dev):
    g = golden("game_mode")
    P = int(g["P"])
    t = np.linspace(-np.pi, np.pi, P)
    curves = dev.host_curve_points(g["v_ecc"], g["v_a"], g["v_b"], g["v_pea"], t)
    scale = np.abs(g["v_curves"]).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(curves - g["v_curves"]) <= 1e-12 * scale)
    moved = dev.host_curve_move_to(g["v_curves"], g["v_body_pos"], g["v_ref"], g["v_f"],
                                   g["v_pea"])
    scale = np.abs(g["v_curves_mov"]).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(moved - g["v_curves_mov"]) <= 1e-12 * scale)


def test_game_body_path_golden(dev):
    """phys_body drop-in on the Solar System: initial(1) + 1000 moves + curve_move."""
    from volatilespace_b200.phys_body import Physics
    g = golden("game_mode")
    pb = Physics(dev, curve_points=int(g["P"]))
    pb.load({"gc": float(g["gc"]), "coi_coef": float(g["coi_coef"])}, {"mass": g["mass"]},
            {"a": g["a"], "ecc": g["ecc"], "pe_arg": g["pea"], "ma": g["ma0"], "ref": g["ref"],
             "dir": g["dr"]})
    body_data, body_orb, pos, ma, ea, cm = pb.initial(1)
    assert set(body_data) >= {"name", "type", "mass", "den", "temp", "lum", "class", "color_b",
                              "color", "radius", "rad_sc", "surf_grav", "atm_h"}
    for k, gk in (("b", "orb_b"), ("f", "orb_f"), ("coi", "orb_coi"), ("n", "orb_n"),
                  ("per", "orb_per"), ("u", "orb_u"), ("pe_d", "orb_pe_d"), ("ap_d", "orb_ap_d")):
        np.testing.assert_allclose(body_orb[k], g[gk], rtol=1e-14, err_msg=k)
    np.testing.assert_allclose(pos, g["i_pos"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(ea, g["i_ea"], rtol=1e-10)
    scale = np.abs(g["i_curves_mov"]).max()
    assert np.abs(cm - g["i_curves_mov"]).max() <= 1e-12 * scale
    for k in range(1, 1001):
        pos, ma, ea = pb.move(3 if k % 2 else 1)
        if k in (1, 1000):
            np.testing.assert_allclose(pos, g["m%d_pos" % k], rtol=1e-10, atol=1e-7)
            np.testing.assert_allclose(ma, g["m%d_ma" % k], rtol=1e-11)
            np.testing.assert_allclose(ea, g["m%d_ea" % k], rtol=1e-10)
            cmv = pb.curve_move()
            assert np.abs(cmv - g["m%d_curves_mov" % k]).max() <= 1e-10 * scale
    one = pb.curve(3)
    assert np.abs(one - g["i_curves"][3]).max() <= 1e-12 * np.abs(g["i_curves"][3]).max()
    vis = pb.curve_move(visible=[2, 5])
    assert set(vis) == {2, 5} and np.array_equal(vis[5], cmv[5])


def test_game_vessel_path_golden(dev):
    """phys_vessel drop-in: 50 moves of 200 mixed elliptic/hyperbolic vessels (incl. ecc 0, 1,
    and the ENRKE quirk region), then curves."""
    from volatilespace_b200.phys_vessel import Physics
    g = golden("game_mode")
    pv = Physics(dev, curve_points=int(g["P"]))
    pv.load({"gc": float(g["gc"])}, {"mass": g["mass"]}, None, None,
            {"a": g["v_a"], "ecc": g["v_ecc"], "pe_arg": g["v_pea"], "ma": g["v_ma0"],
             "ref": g["v_ref"], "dir": g["v_dr"], "ea": g["v_ea0"]})
    from volatilespace_b200.synth import calc_orb
    from volatilespace_b200._lib import KIND_VESSEL
    orb = calc_orb("vessel", pv.ref, pv.body_mass, None, pv.gc, 0.0, pv.a, pv.ecc)
    for k, gk in (("b", "v_b"), ("f", "v_f"), ("n", "v_n"), ("u", "v_u")):
        np.testing.assert_allclose(orb[k], g[gk], rtol=1e-14, err_msg=k)
    pv.b, pv.f, pv.n = orb["b"], orb["f"], orb["n"]
    dev.kepler_load(KIND_VESSEL, pv.curve_points, pv.a, pv.b, pv.f, pv.ecc, pv.pea, pv.n, pv.dr,
                    pv.ma, pv.ea, pv.ref, pv.t)
    for k in range(1, 51):
        pos, ma = pv.move(2, g["v_body_pos"])
        if k in (1, 50):
            np.testing.assert_allclose(ma, g["v%d_ma" % k], rtol=1e-13)
            np.testing.assert_allclose(pv.ea, g["v%d_ea" % k], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(pos, g["v%d_pos" % k], rtol=1e-9, atol=1e-6)
    cm = pv.curve_move()
    scale = np.abs(g["v_curves_mov"]).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(cm - g["v_curves_mov"]) <= 1e-12 * scale)


@pytest.mark.parametrize("n", [200, 5000])
def test_move_state_and_curve_move_equal_the_separate_calls(dev, monkeypatch, n):
    """vsb_kepler_move_state / vsb_kepler_curve_move (move or curves + their results in one call,
    one wait; small states through mapped page-locked memory) against vsb_kepler_move +
    vsb_kepler_get x 3 and vsb_kepler_curves + vsb_kepler_curves_get: bit for bit, on 200 objects
    (mapped path), 5000 (copies) and with the mapped path switched off."""
    from volatilespace_b200 import synth
    from volatilespace_b200._lib import KIND_VESSEL, K_EA, K_MA, K_POS
    k = synth.kepler_objects(n, seed=77)
    P = 16

    def load():
        dev.kepler_load(KIND_VESSEL, P, k["a"], k["b"], k["f"], k["ecc"], k["pea"], k["n"], k["dr"],
                        k["ma"], k["ea"], k["ref"])
    load()
    want = []
    for warp in (1.0, 3.0, 2.0):
        dev.kepler_move(KIND_VESSEL, warp, k["body_pos"])
        dev.kepler_curves(KIND_VESSEL)
        want.append([dev.kepler_get(KIND_VESSEL, f, n) for f in (K_POS, K_MA, K_EA)]
                    + [dev.kepler_curves_get(KIND_VESSEL, 0, n, P)])
    for mapped in (True, False):
        if not mapped:
            monkeypatch.setenv("VSB_NO_MAPPED_FRAME", "1")
        load()
        pos, ma, ea = np.empty((n, 2)), np.empty(n), np.empty(n)
        for step, warp in enumerate((1.0, 3.0, 2.0)):
            dev.kepler_move_state(KIND_VESSEL, warp, k["body_pos"], pos, ma, ea)
            cur = dev.kepler_curve_move(KIND_VESSEL, n, P)
            for got, w in zip((pos, ma, ea, cur), want[step]):
                assert np.array_equal(got, w, equal_nan=True), (mapped, step)
    dev.kepler_free(KIND_VESSEL)


def test_vessel_borrows_resident_body_positions(dev):
    """vessel move with body_pos=None uses the device-resident BODY positions (no host round
    trip): identical to passing the same positions from the host, also after a BODY reload."""
    from volatilespace_b200._lib import KIND_BODY, KIND_VESSEL, K_POS, K_EA
    from volatilespace_b200.phys_body import Physics as BodyPhysics
    from volatilespace_b200.synth import calc_orb
    g = golden("game_mode")
    P = int(g["P"])
    conf = {"gc": float(g["gc"]), "coi_coef": float(g["coi_coef"])}
    orbd = {"a": g["a"], "ecc": g["ecc"], "pe_arg": g["pea"], "ma": g["ma0"], "ref": g["ref"],
            "dir": g["dr"]}
    nv = len(g["v_a"])
    orb = calc_orb("vessel", g["v_ref"], g["mass"], None, conf["gc"], 0.0, g["v_a"], g["v_ecc"])

    def load_vessels():
        dev.kepler_load(KIND_VESSEL, P, g["v_a"], orb["b"], orb["f"], g["v_ecc"], g["v_pea"],
                        orb["n"], g["v_dr"], g["v_ma0"], g["v_ea0"], g["v_ref"])
    res = []
    for mode in ("resident", "host"):
        pb = BodyPhysics(dev, curve_points=P)
        pb.load(conf, {"mass": g["mass"]}, orbd)
        pb.initial(1)
        load_vessels()
        for _ in range(5):
            pos, _, _ = pb.move(2)
            dev.kepler_move(KIND_VESSEL, 2, None if mode == "resident" else pos)
        dev.kepler_curves(KIND_VESSEL)
        res.append((dev.kepler_get(KIND_VESSEL, K_POS, nv), dev.kepler_get(KIND_VESSEL, K_EA, nv),
                    dev.kepler_curves_get(KIND_VESSEL, 0, nv, P)))
        if mode == "resident":      # reload the BODY state while the vessel state borrows it
            pb2 = BodyPhysics(dev, curve_points=P)
            pb2.load(conf, {"mass": g["mass"]}, orbd)
            pb2.initial(1)
            dev.kepler_move(KIND_VESSEL, 0, None)      # re-borrows, must not touch freed memory
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def test_vessel_tick_fused_equals_split_and_oracle(dev, oracle):
    """Config-4 shaped input at a size the oracle finishes in seconds: the fused
    move+curves kernel must equal move followed by curves bit for bit, and match the
    oracle (ragged N, not a multiple of 32)."""
    from volatilespace_b200 import synth
    from volatilespace_b200._lib import KIND_VESSEL, K_EA, K_MA, K_POS
    n, P = 20011, 128
    s = synth.kepler_objects(n, seed=5, quirk=500)
    t = np.linspace(-np.pi, np.pi, P)

    def load():
        dev.kepler_load(KIND_VESSEL, P, s["a"], s["b"], s["f"], s["ecc"], s["pea"], s["n"],
                        s["dr"], s["ma"], s["ea"], s["ref"], t)
    load()
    for _ in range(3):
        dev.kepler_move(KIND_VESSEL, 1.0, s["body_pos"])
    dev.kepler_curves(KIND_VESSEL)
    pos_a = dev.kepler_get(KIND_VESSEL, K_POS, n)
    ea_a = dev.kepler_get(KIND_VESSEL, K_EA, n)
    cur_a = dev.kepler_curves_get(KIND_VESSEL, 0, n, P)
    load()
    for _ in range(3):
        dev.kepler_tick(KIND_VESSEL, 1.0, s["body_pos"])
    assert np.array_equal(pos_a, dev.kepler_get(KIND_VESSEL, K_POS, n))
    assert np.array_equal(ea_a, dev.kepler_get(KIND_VESSEL, K_EA, n))
    assert np.array_equal(cur_a, dev.kepler_curves_get(KIND_VESSEL, 0, n, P))
    ma, ea = s["ma"].copy(), s["ea"].copy()
    for _ in range(3):
        pos, ma, ea = oracle.vessel_move(1.0, s["ref"], s["body_pos"], s["a"], s["b"], s["f"],
                                         s["ecc"], s["pea"], s["n"], s["dr"], ma, ea)
    np.testing.assert_allclose(dev.kepler_get(KIND_VESSEL, K_MA, n), ma, rtol=1e-14)
    np.testing.assert_allclose(ea_a, ea, rtol=1e-9, atol=1e-10)
    rscale = np.maximum(np.abs(s["a"]) * (1 + s["ecc"]), 1.0)[:, None]
    pscale = rscale * np.cosh(np.minimum(np.abs(ea), 20))[:, None]
    assert np.all(np.abs(pos_a - pos) <= 1e-9 * pscale)
    # SURVEY 8d states 1e-12 for ea / pos: met by all but the objects whose solver exit fired one
    # iteration apart (both answers within the solver's 1e-10 of the root)
    assert np.isclose(ea_a, ea, rtol=1e-12, atol=1e-13).mean() > 0.97
    assert (np.abs(pos_a - pos) <= 1e-12 * pscale).all(axis=1).mean() > 0.97
    want = oracle.curve_move_to(oracle.curves_all(s["ecc"], s["a"], s["b"], s["pea"], t),
                                s["body_pos"], s["ref"], s["f"], s["pea"])
    cscale = np.abs(want).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(cur_a - want) <= 1e-12 * cscale)


def test_kepler_full_size_properties(dev):
    """Config 4 at 1/4 size (1M objects, P=512 = 8 GiB of curves): residual of Kepler's
    equation, curve closure and slice consistency -- size-independent properties."""
    from volatilespace_b200 import synth
    from volatilespace_b200._lib import KIND_VESSEL, K_EA, K_MA, K_POS
    n, P = 1 << 20, 512
    s = synth.kepler_objects(n, seed=20240603)
    dev.kepler_load(KIND_VESSEL, P, s["a"], s["b"], s["f"], s["ecc"], s["pea"], s["n"], s["dr"],
                    s["ma"], s["ea"], s["ref"])
    dev.kepler_tick(KIND_VESSEL, 1.0, s["body_pos"])
    ma = dev.kepler_get(KIND_VESSEL, K_MA, n)
    ea = dev.kepler_get(KIND_VESSEL, K_EA, n)
    ell = s["ecc"] < 1
    res_ell = ea[ell] - s["ecc"][ell] * np.sin(ea[ell]) - ma[ell]
    assert np.abs(res_ell).max() < 1e-9
    hyp = np.nonzero(~ell)[0]
    res_hyp = ea[hyp] - s["ecc"][hyp] * np.sinh(ea[hyp]) - ma[hyp]     # negated convention (F8a)
    conv = np.abs(res_hyp) < 1e-8 * np.maximum(1, np.abs(ma[hyp]))
    assert conv.mean() > 0.97
    # the rest is the reference's own behaviour: Newton from ea = ma overshoots for ecc -> 1 and
    # the 50-iteration cap (phys_shared.py:117) ends it early -- the oracle must agree there
    from conftest import ROOT  # noqa: F401
    import sys, os
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    bad = hyp[~conv]
    want = orc.newton_root_kepler_hyp(s["ecc"][bad], ma[bad], s["ea"][bad])
    np.testing.assert_allclose(ea[bad], want, rtol=1e-7, atol=1e-9)
    for first in (0, n // 2 + 17, n - 64):
        cur = dev.kepler_curves_get(KIND_VESSEL, first, 64, P)
        sl = slice(first, first + 64)
        e = s["ecc"][sl] < 1
        # ellipses close on themselves: t = -pi and t = +pi give the same point
        closure = np.abs(cur[e, 0] - cur[e, -1]).max(axis=1)
        assert np.all(closure <= 1e-9 * np.abs(s["a"][sl][e]))
        # every point satisfies the conic equation about its centre
        cp, sp = np.cos(s["pea"][sl]), np.sin(s["pea"][sl])
        cx = s["f"][sl] * cp + s["body_pos"][0, 0]
        cy = s["f"][sl] * sp + s["body_pos"][0, 1]
        x = cur[:, :, 0] - cx[:, None]
        y = cur[:, :, 1] - cy[:, None]
        xr = x * cp[:, None] + y * sp[:, None]
        yr = -x * sp[:, None] + y * cp[:, None]
        sign = np.where(e, 1.0, -1.0)[:, None]
        val = (xr / s["a"][sl][:, None])**2 + sign * (yr / s["b"][sl][:, None])**2
        np.testing.assert_allclose(val[:, ::37], 1.0, rtol=1e-9)
    dev.kepler_free(KIND_VESSEL)


def test_culling_golden(dev):
    """N3: phys_shared.culling mask and the body / vessel visible lists, bit-exact."""
    from volatilespace_b200._lib import KIND_BODY, KIND_VESSEL, K_POS
    g = golden("culling")
    for sb, zoom, mask in zip(g["c_sb"], g["c_zoom"], g["c_mask"]):
        got = dev.host_culling(g["c_coords"], g["c_radius"], sb, float(zoom))
        assert np.array_equal(got, mask)
    from volatilespace_b200.phys_body import Physics as BodyPhysics
    from volatilespace_b200.phys_vessel import Physics as VesselPhysics
    nb, nv = len(g["b_coi"]), len(g["v_ref"])
    z = np.zeros(nb)
    dev.kepler_load(KIND_BODY, 8, g["b_a"], z, z, z, z, z, z, z, z, g["b_ref"])
    dev.kepler_set(KIND_BODY, K_POS, g["b_pos"])
    zv = np.zeros(nv)
    dev.kepler_load(KIND_VESSEL, 8, zv, zv, zv, zv, zv, zv, zv, zv, zv, g["v_ref"])
    dev.kepler_set(KIND_VESSEL, K_POS, g["v_pos"])
    pb = BodyPhysics(dev, curve_points=8)
    pb.n_bodies, pb.coi, pb.a = nb, g["b_coi"], g["b_a"]
    pb.radius, pb.atm_h, pb.screen_x = g["b_radius"], g["b_atm_h"], float(g["b_screen_x"])
    pv = VesselPhysics(dev, curve_points=8)
    pv.n_vessels, pv.pe_d, pv.body_coi = nv, g["v_pe_d"], g["b_coi"]
    pv.screen_x = float(g["b_screen_x"])
    for k, (sb, zoom) in enumerate(zip(g["b_sb"], g["b_zoom"])):
        vb, vc, vo = pb.culling(sb, float(zoom))
        assert np.array_equal(vb, g["b%d_bodies" % k])
        assert np.array_equal(vc, g["b%d_coi" % k])
        assert np.array_equal(vo, g["b%d_orbits" % k])
        vv, vvo = pv.culling(sb, float(zoom), vc)
        assert np.array_equal(vv, g["b%d_vessels" % k])
        assert np.array_equal(vvo, g["b%d_vorbits" % k])


def test_culling_and_gather_at_scale(dev, oracle):
    """1M resident objects: visible list equals the oracle's, gathered curves equal the slices."""
    from volatilespace_b200 import synth
    from volatilespace_b200._lib import KIND_VESSEL, K_POS
    n, P = 1 << 20, 64
    s = synth.kepler_objects(n, seed=77)
    dev.kepler_load(KIND_VESSEL, P, s["a"], s["b"], s["f"], s["ecc"], s["pea"], s["n"], s["dr"],
                    s["ma"], s["ea"], s["ref"])
    dev.kepler_tick(KIND_VESSEL, 1.0, s["body_pos"])
    pos = dev.kepler_get(KIND_VESSEL, K_POS, n)
    sb = np.array([[-2.0e4, 3.0e4], [5.0e4, -1.0e4]])
    rad = np.full(n, 25.0)
    idx = dev.kepler_culling(KIND_VESSEL, n, rad, sb, 0.01)
    want = np.nonzero(oracle.culling(pos, rad, sb, 0.01))[0]
    assert np.array_equal(idx, want) and 100 < len(idx) < n
    assert np.array_equal(dev.kepler_culling(KIND_VESSEL, n, 25.0, sb, 0.01), want)   # one radius for all
    pick = idx[:: max(1, len(idx) // 500)]
    got = dev.kepler_curves_gather(KIND_VESSEL, pick, P)
    for k in (0, len(pick) // 2, len(pick) - 1):
        assert np.array_equal(got[k], dev.kepler_curves_get(KIND_VESSEL, int(pick[k]), 1, P)[0])
    assert len(dev.kepler_culling(KIND_VESSEL, n, None, np.array([[1e9, 2e9], [2e9, 1e9]]), 1.0)) == 0
    dev.kepler_free(KIND_VESSEL)


def test_convert_golden(dev):
    """N1: convert.to_kepler / to_newton against the live reference (index quirks included):
    refs bit-exact, elements and state vectors within 1e-11."""
    from volatilespace_b200 import convert
    g = golden("convert")
    gc, cc = float(g["gc"]), float(g["coi_coef"])
    for tag in ("solar", "c60", "c1500"):
        k = convert.to_kepler(g[tag + "_mass"], {"pos": g[tag + "_pos"], "vel": g[tag + "_vel"]},
                              gc, cc, device=dev)
        assert np.array_equal(k["ref"], g[tag + "_k_ref"]), tag
        for key in ("a", "ecc", "pe_arg", "ma", "dir"):
            np.testing.assert_allclose(k[key], g[tag + "_k_" + key], rtol=1e-11, atol=1e-12,
                                       equal_nan=True, err_msg=tag + key)
    for tag in ("solar", "kn"):
        orb = {key: g[tag + "_k_" + key] for key in ("a", "ecc", "pe_arg", "ma", "ref", "dir")}
        nw = convert.to_newton(g[tag + "_mass"], orb, gc, cc, device=dev)
        np.testing.assert_allclose(nw["pos"], g[tag + "_n_pos"], rtol=1e-10, atol=1e-7)
        np.testing.assert_allclose(nw["vel"], g[tag + "_n_vel"], rtol=1e-10, atol=1e-10)
    orb = {key: g["c60_k_" + key] for key in ("a", "ecc", "pe_arg", "ma", "ref", "dir")}
    with pytest.raises(ValueError):
        convert.to_newton(g["c60_mass"], orb, gc, cc, device=dev)


@pytest.mark.parametrize("mode,rounds,coef", [("scan", "48", 0.7), ("grid", "48", 0.7),
                                              ("grid", "1", 0.7), ("grid", "48", 0.25)])
def test_to_kepler_multi_slab_vs_oracle(dev, oracle, monkeypatch, mode, rounds, coef):
    """to_kepler's sequential first loop, scrambled index order: the blocked sweep over 1024-step
    slabs ("scan"), the fixed-point rounds over the candidate pass ("grid", default from 4096
    bodies), their fall-back to the sweep (round limit 1) and a COI exponent that makes the heavy
    bodies' COIs nest and exceed a grid cell."""
    from volatilespace_b200 import convert
    monkeypatch.setenv("VSB_PARENTS", mode)
    monkeypatch.setenv("VSB_SOC_ROUNDS", rounds)
    rng = np.random.default_rng(99)
    n = 5000
    mass = rng.uniform(0.5, 2.0, n)
    mass[0] = 5.0e4
    mass[1:60] = rng.uniform(100, 3000, 59)
    r = rng.uniform(0.05, 1.0, n) * 6.0e4
    th = rng.uniform(0, 2 * np.pi, n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    sp = np.sqrt(mass[0] / r) * rng.normal(1.0, 0.05, n)
    vel = np.stack([-sp * np.sin(th), sp * np.cos(th)], axis=1)
    pos[0] = 0
    vel[0] = 0
    perm = np.concatenate(([0], 1 + rng.permutation(n - 1)))
    mass, pos, vel = mass[perm], np.ascontiguousarray(pos[perm]), np.ascontiguousarray(vel[perm])
    got = convert.to_kepler(mass, {"pos": pos, "vel": vel}, 1.0, coef, device=dev)
    want = oracle.to_kepler(mass, pos, vel, 1.0, coef)
    assert np.array_equal(got["ref"], want["ref"])
    assert (want["ref"] != 0).sum() > 20
    for key in ("a", "ecc", "pe_arg", "ma", "dir"):
        np.testing.assert_allclose(got[key], want[key], rtol=1e-10, atol=1e-11, equal_nan=True,
                                   err_msg=key)


@pytest.mark.parametrize("mode", ["grid", "scan"])
def test_convert_large_vs_live_reference(dev, monkeypatch, mode):
    """N1 at 4400 bodies against the GENUINE reference (fixture convert_large): the fixed-point
    rounds over the candidate pass (default from 4096 bodies) and the blocked sweep -- refs of all
    bodies bit-exact (index quirks included), elements of the sampled rows within 1e-10."""
    from test_oracle_golden import check_convert_large, convert_large_inputs
    from volatilespace_b200 import convert
    monkeypatch.setenv("VSB_PARENTS", mode)
    mass, pos, vel, g = convert_large_inputs()
    k = convert.to_kepler(mass, {"pos": pos, "vel": vel}, float(g["gc"]), float(g["coi_coef"]), device=dev)
    check_convert_large(k, g, 1e-10)


def test_intersect_kernels_golden(dev):
    """N2: batched solve_quartic / ell_hyp_intersect_circle / predict_enter_coi against the
    live reference's jitted functions."""
    from volatilespace_b200 import orbit_intersect as oi
    g = golden("intersect")
    c = g["q_coef"]
    z = oi.solve_quartic(c[:, 0], c[:, 1], c[:, 2], c[:, 3], c[:, 4], device=dev)
    scale = np.maximum(1.0, np.abs(g["q_roots_re"]) + np.abs(g["q_roots_im"]))
    ok = np.isfinite(g["q_roots_re"]) & np.isfinite(g["q_roots_im"])
    # Ferrari's method takes square roots of differences: where those are ~0 a 1-ulp difference in
    # pow / acos / cos is amplified to ~1e-8, and the first 40 rows are exact biquadratics, the
    # degenerate case of the method (the reference itself returns sqrt(rounding residue) there).
    # In the library's default mode those functions are rounded to nearest like the reference's
    # libm, and the roots -- the degenerate rows included -- agree to 1e-13 of their scale.
    err = np.maximum(np.abs(z.real - g["q_roots_re"]), np.abs(z.imag - g["q_roots_im"]))
    assert np.all(err[ok] <= 1e-13 * scale[ok]), (err[ok] / scale[ok]).max()
    assert np.array_equal(np.isnan(z.real), np.isnan(g["q_roots_re"]))
    ea, cnt = oi.ell_hyp_intersect_circle(g["i_a"], g["i_b"], g["i_ecc"], g["i_cx"], g["i_cy"],
                                          g["i_r"], device=dev)
    assert np.array_equal(cnt, g["i_cnt"])
    # elliptic orbits: rounded-to-nearest atan2 as well; hyperbolic ones end in acosh, where the
    # reference's libm is a 2-ulp routine itself
    ell = g["i_ecc"] < 1
    np.testing.assert_allclose(ea[ell], g["i_ea"][ell], rtol=1e-13, atol=1e-13, equal_nan=True)
    np.testing.assert_allclose(ea[~ell], g["i_ea"][~ell], rtol=1e-11, atol=1e-11, equal_nan=True)
    one = oi.ell_hyp_intersect_circle(g["i_a"][3], g["i_b"][3], g["i_ecc"][3], g["i_cx"][3],
                                      g["i_cy"][3], g["i_r"][3], device=dev)
    assert len(one) == max(1, int(g["i_cnt"][3]))
    out = oi.predict_enter_coi(g["p_vd"], g["p_bd"], 1e-5, 300, device=dev)
    want = g["p_out"]
    assert np.array_equal(np.isnan(out[:, 0]), np.isnan(want[:, 0]))
    hit = ~np.isnan(want[:, 0])
    np.testing.assert_allclose(out[hit], want[hit], rtol=1e-12, atol=1e-12)
    assert hit.sum() >= 20


def test_kepler_errors(dev):
    from volatilespace_b200._lib import VsbError, KIND_BODY, KIND_VESSEL
    dev.kepler_free(KIND_BODY)
    dev.kepler_free(KIND_VESSEL)
    with pytest.raises(VsbError):
        dev.kepler_move(KIND_BODY, 1.0)
    with pytest.raises(VsbError):
        dev.kepler_curves_get(KIND_VESSEL, 0, 1, 8)


# ------------------------------------------------------------- N2 vessel orbit points
def _points_inputs(g):
    nv = len(g["vessel12"])
    nan = lambda w: np.full((nv, w), np.nan)
    return nan(2), nan(2), nan(2), nan(6)


def _assert_points(got, want, what, nan_slack=0, frac=0.0, tol=1e-12):
    """Orbit points against the live-reference fixture in the DEFAULT mode of the library (exact
    libm: the transcendental functions of the quartic and of the march rounded to nearest like the
    reference's glibc): same found / not-found pattern, values to 1e-12 (measured: 2e-16)."""
    flips = np.isnan(got) != np.isnan(want)
    assert flips.any(axis=-1).sum() <= nan_slack, what
    ok = ~np.isnan(want) & ~np.isnan(got)
    np.testing.assert_allclose(got[ok], want[ok], rtol=tol, atol=tol, err_msg=what)
    if frac:
        assert np.isclose(got[ok], want[ok], rtol=1e-11, atol=1e-11).mean() > frac, what


def _robust(v12):
    """Vessels whose COI-entry march is well conditioned in the reference.  predict_enter_coi
    never wraps the mean anomaly (the two wrap statements of orbit_intersect.move are no-ops), so
    over one period it leaves [0, 2 pi]; for e >= 0.8 the ENRKE starter then fails at isolated
    steps and the capped Newton loop returns garbage (|ea| up to 1e7 in the reference itself,
    see DESIGN.md): whether such a step lands inside a COI is decided by the last ulp of sin/cos."""
    ecc = v12[:, 3]
    return ~((ecc >= 0.8) & (ecc < 1))


def test_vessel_points_golden(dev):
    """phys_vessel.Physics.points for all 160 vessels of the live-reference fixture (incl. the
    four whose serial march takes ~20-30 s each in the reference: 110 s there, < 1 s here) and
    the single-vessel variants (entered_coi / left_coi / left_coi_prev_ref / only_enter_coi)."""
    g = golden("vessel_points")
    bi, ba, cl, ce = _points_inputs(g)
    dev.host_vessel_points(g["vessel12"], g["v_ref"], g["body12"], g["body_ref"], bi, ba, cl, ce,
                           steps_limit=int(g["steps_limit"]))
    chunks, rerun, serial, refused = dev.march_stats()
    assert serial == 0 and chunks > 0
    assert rerun <= 0.02 * chunks            # speculation on the hyperbolic chains mostly holds
    # Parity with the live reference, vessel by vessel (measured on the B200, tools/points_diag.py,
    # profiles/r2_points160_diag.txt).  A change in any of these sets is a regression -- or an
    # improvement that must be written down here.
    #   * default mode (exact libm: sin / cos / pow / acos / atan2 of the quartic and of the march
    #     evaluated in double-double arithmetic and rounded to nearest, as the reference's glibc
    #     returns them in > 99.9 % of the cases): every row of every array agrees to 2e-16 relative,
    #     no found / not-found flip, every entered body identical;
    #   * fast mode (CUDA's 1-2 ulp routines, set_march_mode(exact_trig=False)): one flip from the
    #     sign of a rounding residue under the reference's Ferrari sqrt (quartic_solver.py:76-80,
    #     vessel 79), ~1e-8 relative noise on 18 rows from the same degenerate quartic, and five of
    #     the seven high-eccentricity ellipses (0.8 <= e < 1) whose ill-conditioned COI-entry march
    #     (see _robust) hangs on the last bit of sin / cos.
    ill = ~_robust(g["vessel12"])
    want = g["coi_enter"]

    def compare(results, flips_want, off_want, label, elsewhere, loose_max):
        report = []
        for name, got in zip(("body_impact", "body_enter_atm", "coi_leave", "coi_enter"), results):
            ref = g[name]
            flips = set(np.flatnonzero((np.isnan(got) != np.isnan(ref)).any(axis=1)).tolist())
            ok = ~np.isnan(ref) & ~np.isnan(got)
            rel = np.zeros_like(ref)
            rel[ok] = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1.0)
            worst = rel.max(axis=1)
            off = set(np.flatnonzero(worst > 2e-7).tolist())
            loose = int((worst > 1e-11).sum())
            keep = np.array([i not in off for i in range(len(worst))])
            report.append("%s: %d NaN flips %s, %d beyond 2e-7 %s, %d beyond 1e-11, max %.1e elsewhere" % (
                name, len(flips), sorted(flips), len(off), sorted(off), loose, worst[keep].max()))
            assert flips == flips_want[name], report[-1]
            assert off == off_want[name], report[-1]
            assert worst[keep].max() < elsewhere, report[-1]
            assert loose <= loose_max, report[-1]
        print("\nN2 orbit points vs the live reference (160 vessels), %s:\n  " % label + "\n  ".join(report))

    none = {"body_impact": set(), "body_enter_atm": set(), "coi_leave": set(), "coi_enter": set()}
    compare((bi, ba, cl, ce), none, none, "exact libm (default)", elsewhere=1e-14, loose_max=0)
    assert np.array_equal(ce[:, 0], want[:, 0], equal_nan=True)           # entered body, every vessel
    assert (~np.isnan(want[~ill, 0])).sum() >= 12 and (~np.isnan(want[ill, 0])).sum() == 5
    # the fast variant: CUDA's own sin / cos
    dev.set_march_mode(False, exact_trig=False)
    try:
        fast = _points_inputs(g)
        dev.host_vessel_points(g["vessel12"], g["v_ref"], g["body12"], g["body_ref"], *fast,
                               steps_limit=int(g["steps_limit"]))
    finally:
        dev.set_march_mode(False, exact_trig=True)
    # everything else: the biquadratic's sqrt(rounding residue) bounds the agreement at ~1e-8
    compare(fast, dict(none, coi_leave={79}, coi_enter={33, 34}), dict(none, coi_enter={6, 16, 43}),
            "CUDA libm", elsewhere=2e-8, loose_max=12)
    assert all(ill[v] for v in (33, 34, 6, 16, 43))                       # only the e >= 0.8 ellipses
    assert np.array_equal(fast[3][~ill, 0], want[~ill, 0], equal_nan=True)
    opt = lambda x: None if x < 0 else int(x)
    for k, (v, ent, lft, prev, only) in enumerate(g["variants"]):
        bi2, ba2, _, ce2 = _points_inputs(g)
        cl2 = g["var%d_coi_leave_in" % k].copy()
        before = cl2.copy()
        dev.host_vessel_points(g["vessel12"], g["v_ref"], g["body12"], g["body_ref"], bi2, ba2, cl2,
                               ce2, sel=[int(v)], entered_coi=opt(ent), left_coi=opt(lft),
                               left_coi_prev_ref=opt(prev), only_enter_coi=bool(only),
                               steps_limit=int(g["steps_limit"]))
        others = np.arange(len(cl2)) != v
        assert np.array_equal(cl2[others], before[others], equal_nan=True)    # untouched rows
        got = np.concatenate([bi2[v], ba2[v], cl2[v], ce2[v]])
        want_v = g["var%d_rows" % k]
        lo = 4 if only else 0
        _assert_points(got[lo:], want_v[lo:], "variant %d" % k)


def test_march_parallel_is_bit_identical_to_serial(dev):
    """The CTA-parallel COI-entry march (closed-form running sums + speculated Newton chains)
    against the one-thread-per-pair serial loop on the same device: identical bits."""
    rng = np.random.default_rng(77)
    g = golden("vessel_points")
    m = 400
    u = 1.0e4
    hyp = rng.uniform(0, 1, m) < 0.5
    ecc = np.where(hyp, rng.uniform(1.05, 2.5, m), rng.uniform(0, 0.8, m))
    pe = rng.uniform(500, 4000, m)
    a = np.where(hyp, -pe / (ecc - 1), pe / (1 - ecc))
    f = a * ecc
    b = np.sqrt(np.abs(f ** 2 - a ** 2))
    n = np.sqrt(u / np.abs(a) ** 3)
    dr = rng.choice([-1.0, 1.0], m)
    ea = np.where(hyp, -dr * rng.uniform(0.2, 2.0, m), rng.uniform(0, 2 * np.pi, m))
    ma = np.where(hyp, -(ecc * np.sinh(ea) - ea), ea - ecc * np.sin(ea))
    period = np.where(hyp, rng.uniform(2e3, 6e4, m), 2 * np.pi / n)
    period[::17] = rng.uniform(2e5, 6e5, len(period[::17]))      # long marches
    vd = np.stack([a, b, f, ecc, ma, ea, rng.uniform(0, 2 * np.pi, m), period, n, dr], axis=1)
    bhyp = rng.uniform(0, 1, m) < 0.15
    becc = np.where(bhyp, rng.uniform(1.1, 1.6, m), rng.uniform(0, 0.3, m))
    bpe = rng.uniform(800, 5000, m)
    ba_ = np.where(bhyp, -bpe / (becc - 1), bpe / (1 - becc))
    bf = ba_ * becc
    bb = np.sqrt(np.abs(bf ** 2 - ba_ ** 2))
    bn = np.sqrt(u / np.abs(ba_) ** 3)
    bea = np.where(bhyp, rng.uniform(-1.5, 1.5, m), rng.uniform(0, 2 * np.pi, m))
    bma = np.where(bhyp, -(becc * np.sinh(bea) - bea), bea - becc * np.sin(bea))
    coi = rng.uniform(20, 600, m)
    coi[::23] = 0.0
    bd = np.stack([ba_, bb, bf, becc, bma, bea, rng.uniform(0, 2 * np.pi, m), bn,
                   rng.choice([-1.0, 1.0], m), coi], axis=1)
    from volatilespace_b200 import orbit_intersect as oi
    for exact in (False, True):          # fast trigonometry, then the default (exact) mode
        dev.set_march_mode(True, exact_trig=exact)
        try:
            serial = oi.predict_enter_coi(vd, bd, 1e-5, 300, device=dev)
        finally:
            dev.set_march_mode(False, exact_trig=exact)
        par = oi.predict_enter_coi(vd, bd, 1e-5, 300, device=dev)
        chunks, rerun, fell_back, refused = dev.march_stats()
        assert fell_back == 0
        assert np.array_equal(par.view(np.uint64), serial.view(np.uint64)), exact
    hit = ~np.isnan(serial[:, 0])
    assert 20 < hit.sum() < m - 20
    assert rerun <= 0.05 * chunks
    # a march of 5e12 steps is refused (answered "no entry") instead of running for hours
    long_vd, long_bd = vd[:1].copy(), bd[:1].copy()
    long_vd[0, 3], long_vd[0, 7], long_bd[0, 9] = 0.3, 1.0e13, 1.0
    assert np.all(np.isnan(oi.predict_enter_coi(long_vd, long_bd, 1e-5, 300, device=dev)))
    assert dev.march_stats()[3] == 1
    # and on the golden elliptic cases
    gi = golden("intersect")
    dev.set_march_mode(True)
    try:
        s2 = oi.predict_enter_coi(gi["p_vd"], gi["p_bd"], 1e-5, 300, device=dev)
    finally:
        dev.set_march_mode(False)
    p2 = oi.predict_enter_coi(gi["p_vd"], gi["p_bd"], 1e-5, 300, device=dev)
    assert np.array_equal(p2.view(np.uint64), s2.view(np.uint64))


def test_vessel_points_through_the_class(dev):
    """phys_vessel.Physics.load / initial / points with the reference's dict layout."""
    from volatilespace_b200 import phys_vessel
    g = golden("vessel_points")
    v12, b12 = g["vessel12"], g["body12"]
    nb = len(b12)
    sub = np.flatnonzero(v12[:, 3] < 1)[:60]                     # elliptic vessels, quick marches
    body_data = {"mass": np.ones(nb), "radius": b12[:, 10], "atm_h": b12[:, 11]}
    body_orb = {"ref": g["body_ref"], "a": b12[:, 0], "b": b12[:, 1], "f": b12[:, 2],
                "ecc": b12[:, 3], "pea": b12[:, 6], "n": b12[:, 7], "dir": b12[:, 8],
                "coi": b12[:, 9]}
    vorb = {"a": v12[sub, 0], "ecc": v12[sub, 3], "pe_arg": v12[sub, 6], "ma": v12[sub, 4],
            "ref": g["v_ref"][sub], "dir": v12[sub, 8], "ea": v12[sub, 5]}
    pv = phys_vessel.Physics(device=dev, curve_points=8)
    pv.load({"gc": 1.0}, body_data, body_orb, {"name": ["v%d" % i for i in sub]}, vorb)
    # derived elements straight from the fixture (initial() would recompute them from body_mass)
    pv.b, pv.f, pv.n, pv.pe_d, pv.ap_d, pv.period = (v12[sub, k] for k in (1, 2, 7, 9, 10, 11))
    pv.body_ma, pv.body_ea = b12[:, 4].copy(), b12[:, 5].copy()
    pv.predict_coi_limit = int(g["steps_limit"])
    pv.points()
    _assert_points(pv.coi_leave, g["coi_leave"][sub], "coi_leave")
    _assert_points(pv.coi_enter, g["coi_enter"][sub], "coi_enter")     # incl. the e >= 0.8 ellipses
    _assert_points(pv.body_impact, g["body_impact"][sub], "body_impact")
    keep = pv.coi_enter.copy()
    pv.points(vessel=3, only_enter_coi=True)
    assert np.array_equal(pv.coi_enter, keep, equal_nan=True)


# ------------------------------------------------------------- N2 per-tick vessel events
def test_vessel_event_scans_golden(dev):
    """check_warp / predict_enter_coi_service / cross_coi against snapshots of the live reference's
    game loop: decisions (situation, alarm vessel, hold mask, service list, crossing vessel and
    kind, new reference, direction) must be identical; the new elements agree to 1e-9 (they come
    out of atan / acos / acosh chains)."""
    from conftest import VS, events_body7, split_points
    g = golden("vessel_events")
    opt = lambda x: None if x < 0 else int(x)
    for k in range(int(g["n_cw"])):
        v, o = g["cw%d_v" % k], g["cw%d_out" % k]
        bi, ba, cl, ce = split_points(g["cw%d_p" % k])
        hold = g["cw%d_hold_in" % k].copy()
        got = dev.host_check_warp(v[:, VS["ecc"]], v[:, VS["dr"]], v[:, VS["ea"]], v[:, VS["ma"]],
                                  v[:, VS["n"]], bi, ba, cl, ce, o[3], hold)
        assert got == (opt(o[0]), opt(o[1]), int(o[2])), k
        assert np.array_equal(hold, g["cw%d_hold_out" % k]), k
    for k in range(int(g["n_sv"])):
        v, fl = g["sv%d_v" % k], g["sv%d_flags" % k]
        ce = split_points(g["sv%d_p" % k])[3]
        calls = dev.host_enter_coi_service(v[:, VS["dr"]], v[:, VS["prev_ea"]], v[:, VS["ea"]], ce,
                                           fl[0] >= 0 or fl[1] >= 0)
        assert np.array_equal(calls, g["sv%d_calls" % k]), k
    events = 0
    for k in range(int(g["n_cx"])):
        v, ref, ret = g["cx%d_v" % k], g["cx%d_ref" % k], int(g["cx%d_ret" % k][0])
        bi, ba, cl, ce = split_points(g["cx%d_p" % k])
        ves, kind = dev.host_cross_scan(v[:, VS["ecc"]], v[:, VS["dr"]], v[:, VS["prev_ea"]],
                                        v[:, VS["ea"]], bi, cl, ce)
        if ves is None:
            assert ret == -1, k
            continue
        r = int(ref[ves])
        new_ref = int(g["body_ref"][r]) if kind == 2 else int(ce[ves, 0])
        row = v[ves]
        out, status = dev.host_cross_apply(
            kind, [row[0], row[1], row[2], row[3], row[4], row[10], row[6],
                   cl[ves, 0] if kind == 2 else ce[ves, 1]],
            r, new_ref, events_body7(g, g["cx%d_bpos" % k]), float(g["gc"]))
        assert (-1 if status[0] else ves) == ret, k
        if ret >= 0:
            vo = g["cx%d_row_out" % k]
            want = np.array([vo[0], vo[3], vo[4], vo[7], vo[6], g["cx%d_ref_out" % k][0]])
            assert out[0, 4] == want[4] and out[0, 5] == want[5], k       # direction, new reference
            np.testing.assert_allclose(out[0], want, rtol=1e-9, atol=1e-9, err_msg=str(k))
            events += 1
    assert events >= 20
    # edge cases: no vessels, nothing pending
    assert dev.host_check_warp(*([np.zeros(0)] * 5), *[np.zeros((0, 2))] * 3, np.zeros((0, 6)), 1.0,
                               np.zeros(0, dtype=np.uint8)) == (None, None, 0)
    assert dev.host_cross_scan(*([np.zeros(0)] * 4), np.zeros((0, 2)), np.zeros((0, 2)),
                               np.zeros((0, 6))) == (None, 0)


def test_vessel_events_through_the_class(dev):
    """The shim's cross_coi / change_vessel / check_warp / predict_enter_coi_service driven from
    the recorded reference states (attribute-level replay of game.py's tick)."""
    from conftest import VS, split_points
    from volatilespace_b200 import phys_vessel
    g = golden("vessel_events")
    bs = g["body_static"]
    nb = len(bs)
    opt = lambda x: None if x < 0 else int(x)

    def fresh(v, p, ref, flags):
        pv = phys_vessel.Physics(device=dev, curve_points=int(g["P"]))
        body_data = {"mass": g["body_mass"], "radius": bs[:, 8], "atm_h": bs[:, 9]}
        body_orb = {"ref": g["body_ref"], "a": bs[:, 0], "b": bs[:, 1], "f": bs[:, 2], "ecc": bs[:, 3],
                    "pea": bs[:, 4], "n": bs[:, 5], "dir": bs[:, 6], "coi": bs[:, 7]}
        vorb = {"a": v[:, 0].copy(), "ecc": v[:, 3].copy(), "pe_arg": v[:, 4].copy(),
                "ma": v[:, 7].copy(), "ref": ref.copy(), "dir": v[:, 6].copy(), "ea": v[:, 8].copy()}
        pv.load({"gc": float(g["gc"])}, body_data, body_orb,
                {"name": ["v%d" % i for i in range(len(v))]}, vorb)
        pv.b, pv.f, pv.n, pv.u, pv.pe_d, pv.ap_d, pv.period = (v[:, VS[c]].copy() for c in (
            "b", "f", "n", "u", "pe_d", "ap_d", "period"))
        pv.prev_ea = v[:, VS["prev_ea"]].copy()
        pv.body_impact, pv.body_enter_atm, pv.coi_leave, pv.coi_enter = (x.copy() for x in split_points(p))
        pv.entered_coi, pv.left_coi, pv.left_coi_prev_ref = (opt(x) for x in flags)
        pv.predict_coi_limit = int(g["steps_limit"])
        pv.dev.kepler_load(1, pv.curve_points, pv.a, pv.b, pv.f, pv.ecc, pv.pea, pv.n, pv.dr, pv.ma,
                           pv.ea, pv.ref, pv.t)
        return pv

    done = 0
    for k in range(int(g["n_cx"])):
        ret = int(g["cx%d_ret" % k][0])
        pv = fresh(g["cx%d_v" % k], g["cx%d_p" % k], g["cx%d_ref" % k], g["cx%d_flags_in" % k])
        pv.body_pos = g["cx%d_bpos" % k]
        got = pv.cross_coi()
        assert (-1 if got is None else got) == ret, k
        assert [(-1 if x is None else x) for x in (pv.entered_coi, pv.left_coi, pv.left_coi_prev_ref)] \
            == g["cx%d_flags_out" % k].tolist(), k
        if ret >= 0:
            vo = g["cx%d_row_out" % k]
            np.testing.assert_allclose([pv.a[ret], pv.ecc[ret], pv.pea[ret], pv.ma[ret], pv.dr[ret]],
                                       vo[[0, 3, 4, 7, 6]], rtol=1e-9, atol=1e-9)
            assert pv.ref[ret] == g["cx%d_ref_out" % k][0]
            done += 1
    assert done >= 20
    for k in range(0, int(g["n_cv"]), 2):
        ves = int(g["cv%d_vessel" % k][0])
        pv = fresh(g["cv%d_v" % k], g["cv%d_p" % k], g["cv%d_ref" % k], g["cv%d_flags" % k])
        pv.body_ma, pv.body_ea = g["cv%d_bma" % k].copy(), g["cv%d_bea" % k].copy()
        pv.change_vessel(ves)
        vo = g["cv%d_row_out" % k]
        got = np.array([pv.b[ves], pv.f[ves], pv.n[ves], pv.u[ves], pv.pe_d[ves], pv.ap_d[ves],
                        pv.period[ves], pv.ea[ves]])
        np.testing.assert_allclose(got, vo[[1, 2, 5, 10, 11, 12, 13, 8]], rtol=1e-12, atol=1e-12)
        want_p = g["cv%d_p_row_out" % k]
        got_p = np.concatenate([pv.body_impact[ves], pv.body_enter_atm[ves], pv.coi_leave[ves],
                                pv.coi_enter[ves]])
        flips = np.isnan(got_p) != np.isnan(want_p)
        both = ~np.isnan(got_p) & ~np.isnan(want_p)
        assert flips.sum() == 0, k
        np.testing.assert_allclose(got_p[both], want_p[both], rtol=1e-12, atol=1e-12)
    k = 0
    pv = fresh(g["cw%d_v" % k], g["cw%d_p" % k], np.zeros(len(g["cw%d_v" % k]), dtype=np.int64),
               [-1, -1, -1])
    pv._hold[:] = g["cw%d_hold_in" % k]
    o = g["cw%d_out" % k]
    assert pv.check_warp(o[3]) == (opt(o[0]), opt(o[1]), int(o[2]))
    assert np.array_equal(pv.physical_hold, np.flatnonzero(g["cw%d_hold_out" % k]))


def test_vessel_event_argument_errors(dev):
    from volatilespace_b200._lib import VsbError
    z = np.zeros(3)
    with pytest.raises(VsbError):          # event kind outside {1, 2}
        dev.host_cross_apply(3, np.ones(8), 0, 0, np.ones((2, 7)), 1.0)
    with pytest.raises(VsbError):          # body index out of range
        dev.host_cross_apply(2, np.ones(8), 5, 0, np.ones((2, 7)), 1.0)
    nan2, nan6 = np.full((3, 2), np.nan), np.full((3, 6), np.nan)
    with pytest.raises(VsbError):          # vessel reference out of range
        dev.host_vessel_points(np.ones((3, 12)), np.array([0, 9, 0]), np.ones((2, 12)),
                               np.array([0, 0]), nan2.copy(), nan2.copy(), nan2.copy(), nan6.copy())
    with pytest.raises(VsbError):          # selected vessel out of range
        dev.host_vessel_points(np.ones((3, 12)), np.zeros(3, dtype=np.int64), np.ones((2, 12)),
                               np.array([0, 0]), nan2.copy(), nan2.copy(), nan2.copy(), nan6.copy(),
                               sel=[7])
    # entering the body the vessel already orbits: the reference returns None
    out, status = dev.host_cross_apply(1, np.ones(8), 1, 1, np.ones((2, 7)), 1.0)
    assert status[0] == 1 and np.all(np.isnan(out))
    # all-NaN points: nothing to do, holds are released
    hold = np.ones(3, dtype=np.uint8)
    assert dev.host_check_warp(z + 0.5, z + 1, z, z, z + 0.01, nan2, nan2, nan2, nan6, 1.0, hold) \
        == (None, None, 0)
    assert not hold.any()
    # curve_segments: needs the resident event state and generated curves, indices in range
    from volatilespace_b200._lib import KIND_VESSEL
    dev.kepler_free(KIND_VESSEL)
    with pytest.raises(VsbError):          # nothing loaded
        dev.vessel_curve_segments([0], 8)
    one = np.ones(4)
    dev.kepler_load(KIND_VESSEL, 8, one * 100, one * 90, one * 40, one * 0.4, one, one * 0.01, one,
                    one, one, np.zeros(4, dtype=np.int64), np.linspace(-np.pi, np.pi, 8))
    with pytest.raises(VsbError):          # no event state yet
        dev.vessel_curve_segments([0], 8)
    dev.vessel_events_load(one, one, one)
    with pytest.raises(VsbError):          # no curves / parent positions yet
        dev.vessel_curve_segments([0], 8)
    dev.kepler_tick(KIND_VESSEL, 1.0, np.zeros((1, 2)))
    with pytest.raises(VsbError):          # vessel index out of range
        dev.vessel_curve_segments([4], 8)
    light, dark, first, itype, itime, srange = dev.vessel_curve_segments([3, 1], 8)
    assert itype.tolist() == [0, 0] and np.all(np.isnan(dark)) and not np.isnan(light).any()
    assert np.array_equal(light[0], dev.kepler_curves_get(KIND_VESSEL, 3, 1, 8)[0])
    assert light.shape == (2, 8, 2) and dev.vessel_curve_segments([], 8)[0].shape == (0, 8, 2)
    dev.kepler_free(KIND_VESSEL)


def test_resident_vessel_events(dev):
    """The device-resident form of the vessel event state: same kernels as the host-array calls,
    so the answers must be identical bit for bit; kepler_move keeps prev_ea."""
    from conftest import VS, events_body12, events_vessel12, split_points
    from volatilespace_b200._lib import KIND_VESSEL, K_EA
    g = golden("vessel_events")
    opt = lambda x: None if x < 0 else int(x)
    P = 8
    t = np.linspace(-np.pi, np.pi, P)

    def load(v, p, ref, hold=None):
        col = lambda name: np.ascontiguousarray(v[:, VS[name]])
        dev.kepler_load(KIND_VESSEL, P, col("a"), col("b"), col("f"), col("ecc"), col("pea"),
                        col("n"), col("dr"), col("ma"), col("ea"), ref, t)
        bi, ba, cl, ce = split_points(p)
        dev.vessel_events_load(col("pe_d"), col("ap_d"), col("period"), bi, ba, cl, ce, hold,
                               col("prev_ea"))
        return bi, ba, cl, ce

    nv = len(g["cw0_v"])
    zref = np.zeros(nv, dtype=np.int64)
    for k in range(0, int(g["n_cw"]), 3):
        v, o = g["cw%d_v" % k], g["cw%d_out" % k]
        load(v, g["cw%d_p" % k], zref, g["cw%d_hold_in" % k])
        assert dev.vessel_check_warp(o[3]) == (opt(o[0]), opt(o[1]), int(o[2])), k
        assert np.array_equal(dev.vessel_events_get(nv)["hold"], g["cw%d_hold_out" % k]), k
    for k in range(0, int(g["n_sv"]), 2):
        v, fl = g["sv%d_v" % k], g["sv%d_flags" % k]
        load(v, g["sv%d_p" % k], zref)
        calls = dev.vessel_enter_coi_service(nv, fl[0] >= 0 or fl[1] >= 0)
        assert np.array_equal(calls, g["sv%d_calls" % k]), k
    for k in range(0, int(g["n_cx"]), 2):
        v = g["cx%d_v" % k]
        bi, ba, cl, ce = load(v, g["cx%d_p" % k], g["cx%d_ref" % k])
        want = dev.host_cross_scan(v[:, VS["ecc"]], v[:, VS["dr"]], v[:, VS["prev_ea"]],
                                   v[:, VS["ea"]], bi, cl, ce)
        assert dev.vessel_cross_scan() == want, k
    # points on resident state == points over host arrays (one change_vessel snapshot)
    k = 0
    v, ref, fl = g["cv%d_v" % k], g["cv%d_ref" % k], g["cv%d_flags" % k]
    sel = np.flatnonzero(v[:, VS["ecc"]] < 0.8)[:50]
    bi, ba, cl, ce = (x.copy() for x in load(v, g["cv%d_p" % k], ref))
    body12 = events_body12(g, g["cv%d_bma" % k], g["cv%d_bea" % k])
    kw = dict(sel=sel, entered_coi=opt(fl[0]), left_coi=opt(fl[1]), left_coi_prev_ref=opt(fl[2]),
              steps_limit=int(g["steps_limit"]))
    dev.vessel_points(body12, g["body_ref"], **kw)
    dev.host_vessel_points(events_vessel12(v), ref, body12, g["body_ref"], bi, ba, cl, ce, **kw)
    got = dev.vessel_events_get(nv)
    for key, want in (("body_impact", bi), ("body_enter_atm", ba), ("coi_leave", cl),
                      ("coi_enter", ce)):
        assert np.array_equal(got[key], want, equal_nan=True), key
    # kepler_move keeps the previous anomalies
    ea_before = dev.kepler_get(KIND_VESSEL, K_EA, nv)
    dev.kepler_move(KIND_VESSEL, 3.0, np.zeros((len(g["body_mass"]), 2)))
    assert np.array_equal(dev.vessel_events_get(nv)["prev_ea"], ea_before)
    assert not np.array_equal(dev.kepler_get(KIND_VESSEL, K_EA, nv), ea_before)
    dev.vessel_events_set_orbit(5, 1.0, 2.0, 3.0)
    from volatilespace_b200._lib import VsbError
    dev.kepler_free(KIND_VESSEL)
    with pytest.raises(VsbError):
        dev.vessel_check_warp(1.0)


def test_game_loop_resident_equals_host_arrays():
    """game.py's physics tick (check_warp, predict_enter_coi_service, move, cross_coi,
    change_vessel) driven through phys_vessel.Physics twice -- event state resident on the device
    vs shipped with every call -- must evolve identically, tick by tick."""
    from conftest import VS
    from volatilespace_b200 import phys_vessel
    from volatilespace_b200.device import Device
    g, gp = golden("vessel_events"), golden("vessel_points")
    bs = g["body_static"]
    keep = g["keep"]
    v12 = gp["vessel12"][keep]
    body_data = {"mass": g["body_mass"], "radius": bs[:, 8], "atm_h": bs[:, 9]}
    body_orb = {"ref": g["body_ref"], "a": bs[:, 0], "b": bs[:, 1], "f": bs[:, 2], "ecc": bs[:, 3],
                "pea": bs[:, 4], "n": bs[:, 5], "dir": bs[:, 6], "coi": bs[:, 7]}
    bpos = g["cx0_bpos"]
    bma, bea = gp["body12"][:, 4].copy(), gp["body12"][:, 5].copy()
    sims = []
    for resident in (True, False):
        pv = phys_vessel.Physics(device=Device(0), curve_points=8)
        pv.resident_events = resident
        vorb = {"a": v12[:, 0].copy(), "ecc": v12[:, 3].copy(), "pe_arg": v12[:, 6].copy(),
                "ma": v12[:, 4].copy(), "ref": gp["v_ref"][keep].copy(), "dir": v12[:, 8].copy()}
        pv.load({"gc": float(g["gc"])}, body_data, body_orb,
                {"name": ["v%d" % i for i in keep]}, vorb)
        pv.predict_coi_limit = int(g["steps_limit"])
        pv.initial(1, bpos, bma, bea)
        sims.append(pv)
    a, b = sims
    assert a._resident and not b._resident

    def same(what):
        for key in ("a", "ecc", "pea", "ma", "ea", "dr", "ref", "pos", "body_impact",
                    "body_enter_atm", "coi_leave", "coi_enter", "pe_d", "period"):
            assert np.array_equal(getattr(a, key), getattr(b, key), equal_nan=True), (what, key)
        assert (a.entered_coi, a.left_coi, a.left_coi_prev_ref) == \
            (b.entered_coi, b.left_coi, b.left_coi_prev_ref), what
        assert np.array_equal(a.physical_hold, b.physical_hold), what

    same("initial")
    assert (~np.isnan(a.coi_leave[:, 0])).sum() > 10
    crossings = alarms = 0
    for tick in range(400):
        warp = (1, 5, 25, 100)[(tick // 100) % 4]
        ra, rb = a.check_warp(warp), b.check_warp(warp)
        assert ra == rb, tick
        alarms += ra[0] is not None
        assert np.array_equal(a.predict_enter_coi_service(), b.predict_enter_coi_service()), tick
        a.move(warp, bpos, bma, bea)
        b.move(warp, bpos, bma, bea)
        ca, cb = a.cross_coi(), b.cross_coi()
        assert ca == cb, tick
        if ca is not None:
            crossings += 1
            a.change_vessel(ca)
            b.change_vessel(cb)
        if tick % 20 == 0 or ca is not None:
            same("tick %d" % tick)
    same("end")
    assert crossings >= 5 and alarms >= 5
    for pv in sims:
        pv.dev.close()


def test_convert_velocity_helpers_golden(dev):
    """Batched convert.kepler_to_velocity / velocity_to_kepler and phys_shared.orb2xy against the
    live reference: branch decisions (direction, failsafe flips) identical, values to 1e-9."""
    from volatilespace_b200 import convert
    from volatilespace_b200._lib import check, f64, ptr
    g = golden("convert_vel")
    n = len(g["a"])
    xy = np.empty((n, 2))
    check(dev.L.vsb_host_orb2xy(dev.h, n, *(ptr(f64(g[k])) for k in ("a", "b", "f", "ecc", "pea",
                                                                     "ref_pos", "ea")), ptr(xy)))
    np.testing.assert_allclose(xy, g["xy"], rtol=1e-12, atol=1e-9)
    assert np.all(xy[:3] == 0.0)                     # ea == 0: (0, 0) without the reference position
    vel = convert.kepler_to_velocity(g["rel"], g["a"], g["ecc"], g["pea"], g["u"], g["dr"], device=dev)
    ok = np.isfinite(g["vel"]).all(axis=1)
    assert np.array_equal(np.isfinite(vel).all(axis=1), ok)
    np.testing.assert_allclose(vel[ok], g["vel"][ok], rtol=1e-9, atol=1e-9)
    one = convert.kepler_to_velocity(g["rel"][7], g["a"][7], g["ecc"][7], g["pea"][7], g["u"][7],
                                     g["dr"][7], device=dev)
    assert np.array_equal(one, vel[7])
    for fs in (0, 1, 2):
        want = g["v2k_fs%d" % fs]
        vin = g["v2k_vel"].copy()
        got = convert.velocity_to_kepler(g["v2k_pos"], vin, g["v2k_u"], fs, device=dev)
        assert np.array_equal(vin, g["v2k_vel"])     # caller's array untouched
        assert np.array_equal(got[:, 4], want[:, 4]), fs                  # direction
        close = np.isclose(got, want, rtol=1e-9, atol=1e-9, equal_nan=True).all(axis=1)
        assert close.mean() > 0.99, fs               # a failsafe decision may sit on the 1 % edge
    a, ecc, pe, ma, dr = convert.velocity_to_kepler(g["v2k_pos"][0], g["v2k_vel"][0], g["v2k_u"][0],
                                                    1, device=dev)
    assert isinstance(dr, int) and np.isclose(a, g["v2k_fs1"][0, 0], rtol=1e-9)


def test_editor_free_functions_golden(dev):
    """phys_editor.{find_parent_one, gravity_one, kepler_basic_one, check_collision_one} of the
    shim against the live reference's jitted functions (pair_kernels fixture)."""
    from volatilespace_b200 import phys_editor as pe
    g = golden("pair_kernels")
    n = len(g["fp_mass"])
    parents = pe.find_parents_all(g["fp_order"], g["fp_pos"], g["fp_coi"], device=dev)
    assert np.array_equal(parents, g["fp_parents"])
    for body in (0, 1, 57, n - 1):
        assert pe.find_parent_one(body, g["fp_order"], g["fp_pos"], g["fp_coi"], device=dev) == \
            g["fp_parents"][body]
    rel_pos = g["fp_pos"][parents] - g["fp_pos"]
    for body in (0, 3, 200):
        acc = pe.gravity_one(body, int(parents[body]), g["fp_mass"], rel_pos[body], float(g["g1_gc"]),
                             device=dev)
        np.testing.assert_allclose(acc, g["g1_acc"][body], rtol=1e-12, atol=1e-300)
    f, ev, a, b, w, c = pe.kepler_basic_all(parents, g["fp_mass"], g["fp_pos"], g["kb_vel"],
                                            float(g["g1_gc"]), 0.7, device=dev)
    for got, key in ((f, "kb_focus"), (ev, "kb_ecc_v"), (a, "kb_semi_major"), (b, "kb_semi_minor"),
                     (w, "kb_periapsis_arg"), (c, "kb_coi")):
        np.testing.assert_allclose(got, g[key], rtol=1e-10, atol=1e-10, err_msg=key)
    one = pe.kepler_basic_one(5, int(parents[5]), g["fp_mass"], g["fp_pos"], g["kb_vel"],
                              float(g["g1_gc"]), 0.7, device=dev)
    np.testing.assert_allclose(one[2], g["kb_semi_major"][5], rtol=1e-10)
    np.testing.assert_allclose(one[1], g["kb_ecc_v"][5], rtol=1e-10, atol=1e-12)
    # collisions: every case of the fixture through the per-body form
    cm = g["cc_mass"]
    for k in range(len(g["cc_del"])):
        pos, rad = g["cc_pos"][k], g["cc_rad"][k]
        d, a_ = int(g["cc_del"][k]), int(g["cc_add"][k])
        if d < 0:
            assert pe.check_collision_one(0, cm, pos, rad, device=dev) is None
            continue
        i = min(d, a_)
        assert pe.check_collision_one(i, cm, pos, rad, device=dev) == (d, a_)
        if i > 0:                                    # an earlier body collides with nobody
            assert pe.check_collision_one(i - 1, cm, pos, rad, device=dev) is None
    assert pe.check_collision_one(len(cm) - 1, cm, g["cc_pos"][0], g["cc_rad"][0], device=dev) is None


# ------------------------------------------------------------------ curve_segments
@pytest.mark.gpu
def test_curve_segments_golden(dev):
    """phys_vessel.Physics.curve_segments on the resident state against the live reference:
    intersection types, NaN pattern (= which rows each part holds) and select ranges exact;
    coordinates to 1e-9 of the curve's scale (the curves are generated on the device)."""
    from conftest import VS, split_points
    from volatilespace_b200._lib import KIND_VESSEL, K_EA, K_MA, K_POS
    g = golden("vessel_draw")
    P, nv = int(g["P"]), len(g["keep"])
    t = np.linspace(-np.pi, np.pi, P)
    for k in range(int(g["n_snap"])):
        v, p, ref, pos, bpos = (g["s%d_%s" % (k, x)] for x in ("v", "p", "ref", "pos", "bpos"))
        col = lambda name: np.ascontiguousarray(v[:, VS[name]])
        dev.kepler_load(KIND_VESSEL, P, col("a"), col("b"), col("f"), col("ecc"), col("pea"),
                        col("n"), col("dr"), col("ma"), col("ea"), ref, t)
        dev.vessel_events_load(col("pe_d"), col("ap_d"), col("period"), *split_points(p), None,
                               col("prev_ea"))
        dev.kepler_move(KIND_VESSEL, 0.0, bpos)          # parent positions onto the device
        dev.kepler_set(KIND_VESSEL, K_EA, col("ea"))     # ... and the reference's exact state
        dev.kepler_set(KIND_VESSEL, K_MA, col("ma"))
        dev.kepler_set(KIND_VESSEL, K_POS, pos)
        dev.kepler_curves(KIND_VESSEL)
        cm = dev.kepler_curves_get(KIND_VESSEL, 0, nv, P)
        scale = np.max(np.abs(g["s%d_curves_mov" % k]), axis=(1, 2))
        assert np.all(np.abs(cm - g["s%d_curves_mov" % k]) <= 1e-12 * scale[:, None, None])
        idx = np.arange(nv)
        light, dark, first, itype, itime, srange = dev.vessel_curve_segments(idx, P)
        assert np.array_equal(itype, g["s%d_itype" % k]), k
        assert np.array_equal(srange, g["s%d_srange" % k], equal_nan=True), k
        for got, want in ((light, g["s%d_light" % k]), (dark, g["s%d_dark" % k])):
            assert np.array_equal(np.isnan(got), np.isnan(want)), k
            err = np.abs(np.nan_to_num(got) - np.nan_to_num(want))
            assert np.all(err <= 1e-9 * scale[:, None, None]), (k, float(err.max()))
        assert np.array_equal(np.isnan(first), np.isnan(g["s%d_first" % k]))
        assert np.all(np.abs(np.nan_to_num(first) - np.nan_to_num(g["s%d_first" % k]))
                      <= 1e-9 * scale[:, None])
        np.testing.assert_allclose(itime, g["s%d_itime" % k], rtol=1e-9, atol=1e-9)
        # a subset in another order gives the same rows
        sub = np.array([nv - 1, 3, 77, 0])
        l2, d2, f2, t2, tm2, r2 = dev.vessel_curve_segments(sub, P)
        assert np.array_equal(l2, light[sub], equal_nan=True)
        assert np.array_equal(d2, dark[sub], equal_nan=True) and np.array_equal(t2, itype[sub])
    dev.kepler_free(KIND_VESSEL)


@pytest.mark.gpu
def test_curve_segments_through_the_class():
    """The shim method: full-shaped arrays like the reference, NaN rows outside visible_orbits."""
    from conftest import VS, split_points
    from volatilespace_b200.phys_vessel import Physics
    g = golden("vessel_draw")
    P, nv = int(g["P"]), len(g["keep"])
    k = 1
    v, p, ref, pos, bpos = (g["s%d_%s" % (k, x)] for x in ("v", "p", "ref", "pos", "bpos"))
    nb = len(g["body_mass"])
    ph = Physics(curve_points=P)
    ph.load({"gc": float(g["gc"])}, {"mass": g["body_mass"]}, {"coi": g["body_static"][:, 7]},
            {"name": ["v%d" % i for i in range(nv)]},
            {"a": v[:, VS["a"]], "ecc": v[:, VS["ecc"]], "pe_arg": v[:, VS["pea"]],
             "ma": v[:, VS["ma"]], "ref": ref, "dir": v[:, VS["dr"]], "ea": v[:, VS["ea"]]})
    ph.initial(0.0, bpos)
    bi, ba, cl, ce = split_points(p)
    ph.body_impact[:], ph.body_enter_atm[:], ph.coi_leave[:], ph.coi_enter[:] = bi, ba, cl, ce
    ph.dev.vessel_events_load(ph.pe_d, ph.ap_d, ph.period, bi, ba, cl, ce, None, v[:, VS["prev_ea"]])
    ph.visible_orbits = np.arange(0, nv, 2)
    light, dark, first, itype, itime, srange = ph.curve_segments()
    assert light.shape == (nv, P, 2) and np.all(np.isnan(light[1::2])) and np.all(itype[1::2] == 0)
    assert np.array_equal(itype[::2], g["s%d_itype" % k][::2])
    scale = np.max(np.abs(g["s%d_curves_mov" % k]), axis=(1, 2))[::2, None, None]
    want = g["s%d_light" % k][::2]
    assert np.array_equal(np.isnan(light[::2]), np.isnan(want))
    assert np.all(np.abs(np.nan_to_num(light[::2]) - np.nan_to_num(want)) <= 1e-7 * scale)
    sparse = ph.curve_segments(dense=False)
    assert sorted(sparse[0]) == list(range(0, nv, 2))
    assert np.array_equal(sparse[0][4], light[4], equal_nan=True)
    assert np.array_equal(ph.intersect_time, itime)


@pytest.mark.gpu
def test_predict_next_orbit_through_the_class():
    """phys_vessel.Physics.predict_next_orbit of the shim against the live reference: the same
    vessels get an orbit, new parent / enter flag exact, scalars rel 1e-7, curve rows 1e-7 of the
    curve's scale with the same NaN pattern (index rounding can move one row on a few)."""
    from conftest import VS, split_points
    from volatilespace_b200.phys_vessel import Physics
    g = golden("vessel_draw")
    P, nv = int(g["P"]), len(g["keep"])
    bs = g["body_static"]
    total = pattern_ok = 0
    for k in range(int(g["n_snap"])):
        v, p, ref, bpos = (g["s%d_%s" % (k, x)] for x in ("v", "p", "ref", "bpos"))
        ph = Physics(curve_points=P)
        ph.load({"gc": float(g["gc"])}, {"mass": g["body_mass"]},
                {"coi": bs[:, 7], "ref": g["body_ref"], "a": bs[:, 0], "b": bs[:, 1], "f": bs[:, 2],
                 "ecc": bs[:, 3], "pea": bs[:, 4], "n": bs[:, 5], "dir": bs[:, 6]},
                {"name": ["v%d" % i for i in range(nv)]},
                {"a": v[:, VS["a"]], "ecc": v[:, VS["ecc"]], "pe_arg": v[:, VS["pea"]],
                 "ma": v[:, VS["ma"]], "ref": ref, "dir": v[:, VS["dr"]], "ea": v[:, VS["ea"]]})
        ph.initial(0.0, bpos)
        bi, ba, cl, ce = split_points(p)
        ph.body_impact[:], ph.body_enter_atm[:], ph.coi_leave[:], ph.coi_enter[:] = bi, ba, cl, ce
        ph.body_ma = g["s%d_bma" % k]
        ph.intersect_time = g["s%d_itime" % k]
        want_v = list(g["s%d_pno_v" % k])
        for i in range(nv):
            r = ph.predict_next_orbit(i)
            if i not in want_v:
                continue
            j = want_v.index(i)
            assert r[0] is not None, (k, i)
            curve, ap, ap_d, ap_t, pe, pe_d, pe_t, new_ref, nrp, nvp, enter = r
            sc = g["s%d_pno_sc" % k][j]
            assert new_ref == int(sc[8]) and bool(enter) == bool(sc[13]), (k, i)
            got = np.array([ap[0], ap[1], ap_d, ap_t, pe[0], pe[1], pe_d, pe_t, nrp[0], nrp[1],
                            nvp[0], nvp[1]], dtype=float)
            want = np.concatenate([sc[:8], sc[9:13]])
            scale = max(np.max(np.abs(want)), 1.0)
            assert np.all(np.abs(got - want) <= 1e-7 * scale), (k, i, got - want)
            wc = g["s%d_pno_curve" % k][j]
            total += 1
            if np.array_equal(np.isnan(curve), np.isnan(wc)):
                pattern_ok += 1
                cs = np.nanmax(np.abs(wc))
                assert np.all(np.abs(np.nan_to_num(curve) - np.nan_to_num(wc)) <= 1e-7 * cs), (k, i)
        ph.dev.kepler_free(1)
    assert total > 100 and pattern_ok >= total - 3
