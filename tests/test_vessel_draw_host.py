"""Host-side drawing helpers of the vessel shim (volatilespace_b200/vessel_draw.py) against the
live-reference fixture tests/golden/vessel_draw.npz.  CPU only: `selected`, `rotate` and
`concat_wrap` are plain host code (same libm as the reference -> bit-exact)."""
import numpy as np

from conftest import VS, golden


def test_selected_and_rotate_bit_exact():
    from volatilespace_b200 import vessel_draw as vd
    g = golden("vessel_draw")
    names = ("a", "b", "ecc", "pea", "n", "dr", "ma", "ea", "u", "pe_d", "ap_d")
    checked = 0
    for k in range(int(g["n_snap"])):
        v, ref, pos, bpos = (g["s%d_%s" % (k, x)] for x in ("v", "ref", "pos", "bpos"))
        sel = g["s%d_selected" % k]
        for i in range(len(v)):
            if np.isnan(sel[i, 0]):
                continue
            el = {n: float(v[i, VS[n]]) for n in names}
            ta, pe, pe_t, ap, ap_t, d, so, sh, sv = vd.selected(el, pos[i], bpos[ref[i]])
            got = np.array([ta, pe[0], pe[1], pe_t, ap[0], ap[1], ap_t, d, so, sh, sv], dtype=float)
            assert np.array_equal(got, sel[i, :11]), (k, i)
            checked += 1
        acc = g["rot_acc"]
        ang, spd = g["s%d_rot_angle_in" % k].copy(), g["s%d_rot_speed_in" % k].copy()
        ang = vd.rotate(1, 3, 1.0, ang, spd, acc)
        assert np.array_equal(ang, g["s%d_rot_angle1" % k])
        ang = vd.rotate(1, 3, -1.0, ang, spd, acc)
        assert np.array_equal(ang, g["s%d_rot_angle2" % k]) and np.array_equal(spd, g["s%d_rot_speed2" % k])
        ang = vd.rotate(5, 3, 1.0, ang, spd, acc)
        assert np.array_equal(ang, g["s%d_rot_angle3" % k]) and np.array_equal(spd, g["s%d_rot_speed3" % k])
    assert checked == 420


def test_concat_wrap_matches_reference_rows():
    """Every light / dark row of the fixture that is a cut of its curve is reproduced by the host
    concat_wrap from the row's own end points and run (normal, wrapped and NaN-cleaned forms)."""
    from volatilespace_b200 import vessel_draw as vd
    rng = np.random.default_rng(0)
    arr = rng.uniform(-1, 1, (12, 2))
    a, b = np.array([9.0, 9.0]), np.array([8.0, 8.0])
    out = vd.concat_wrap(arr, a, 3, 7, b)
    assert np.array_equal(out[:6], np.vstack([a, arr[3:7], b])) and np.all(np.isnan(out[6:]))
    out = vd.concat_wrap(arr, a, 9, 2, b)
    assert np.array_equal(out[:7], np.vstack([b, arr[9:], arr[:2], a])) and np.all(np.isnan(out[7:]))
    out = vd.concat_wrap(arr, a, 0, 12, b)          # truncated to fit
    assert np.array_equal(out, np.vstack([a, arr[:10], b]))
    out = vd.concat_wrap(arr, a, 5, -1, b)          # numpy's negative stop: all but the last row
    assert np.array_equal(out[0], b) and np.array_equal(out[1:8], arr[5:]) and np.array_equal(out[8:11], arr[:3])
    arr[[2, 7]] = np.nan
    out = vd.concat_wrap(arr, a, 5, 10, b)
    want = np.vstack([a, arr[5:7], arr[8:10], b])
    assert np.array_equal(out[:len(want)], want) and np.all(np.isnan(out[len(want):]))
