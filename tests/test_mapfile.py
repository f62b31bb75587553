"""SURVEY 8f N4 (savefile bulk I/O): volatilespace_b200.mapfile against files written and read
back by the reference's own peripherals.save_file / load_file (tests/golden/maps/*.ini,
mapfiles.npz; oracle/gen_golden.py mapfiles)."""
import os
import time

import numpy as np
import pytest

from conftest import golden

MAPS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "maps")


def _check_loaded(res, g, tag):
    gd, cf, b, bo, v, vo = res
    want = g[tag + "_game"]
    assert [gd["name"], gd["date"], str(gd["time"]), str(gd["vessel"])] == list(want)
    assert list(cf.keys()) == list(g[tag + "_conf_keys"])
    assert np.array_equal(np.array([float(x) for x in cf.values()]), g[tag + "_conf_vals"])
    for prefix, d in (("_b_", b), ("_bo_", bo), ("_v_", v), ("_vo_", vo)):
        keys = [k[len(tag + prefix):] for k in g.files if k.startswith(tag + prefix)]
        assert sorted(keys) == sorted(d.keys()), (tag, prefix)
        for k in keys:
            got, ref = np.asarray(d[k]), g[tag + prefix + k]
            assert got.shape == ref.shape, (tag, k, got.shape, ref.shape)
            assert got.dtype.kind == ref.dtype.kind, (tag, k)
            assert np.array_equal(got, ref), (tag, k)


@pytest.mark.parametrize("tag", ["newton", "kepler"])
def test_load_matches_reference(tag):
    from volatilespace_b200 import mapfile
    _check_loaded(mapfile.load_file(os.path.join(MAPS, tag + ".ini")), golden("mapfiles"), tag)


@pytest.mark.parametrize("tag", ["newton", "kepler"])
def test_save_writes_the_reference_bytes(tag, tmp_path):
    """load -> save reproduces the file the reference wrote, byte for byte (incl. the renamed
    duplicate section and the omitted zero atmospheres)."""
    from volatilespace_b200 import mapfile
    src = os.path.join(MAPS, tag + ".ini")
    gd, cf, b, bo, v, vo = mapfile.load_file(src)
    bo = {k: a for k, a in bo.items() if k != "kepler"}
    out = tmp_path / "again.ini"
    # the reference was given the un-renamed names; loading returns the renamed one
    b["name"][3] = b["name"][2]
    mapfile.save_file(str(out), gd, cf, b, bo, v, vo)
    assert out.read_text() == open(src).read()
    _check_loaded(mapfile.load_file(str(out)), golden("mapfiles"), tag)


def test_large_map_round_trip_is_linear(tmp_path):
    """200 000 bodies written and read back in seconds (the reference's per-body np.append makes
    loading quadratic: 24 s at 40 000 bodies in the dev container)."""
    from volatilespace_b200 import mapfile
    rng = np.random.default_rng(1)
    n = 200_000
    b = {"name": np.array(["b%d" % i for i in range(n)]), "mass": rng.uniform(1, 9, n),
         "den": rng.uniform(1e3, 5e3, n), "color": rng.integers(0, 256, (n, 3)),
         "atm_pres0": np.zeros(n), "atm_scale_h": np.zeros(n), "atm_den0": np.zeros(n)}
    bo = {"pos": rng.uniform(-1e6, 1e6, (n, 2)), "vel": rng.uniform(-1, 1, (n, 2))}
    gd = {"name": "big", "date": "x", "time": 0.0, "vessel": None}
    path = str(tmp_path / "big.ini")
    t0 = time.perf_counter()
    mapfile.save_file(path, gd, dict(mapfile.SIM_CONFIG_DEFAULTS), b, bo)
    res = mapfile.load_file(path)
    assert time.perf_counter() - t0 < 60
    assert np.array_equal(res[2]["mass"], b["mass"]) and np.array_equal(res[3]["pos"], bo["pos"])
    assert np.array_equal(res[3]["vel"], bo["vel"]) and np.array_equal(res[2]["color"], b["color"])
    assert res[0]["vessel"] is None and not res[3]["kepler"]


def test_errors(tmp_path):
    from volatilespace_b200 import mapfile
    p = tmp_path / "dup.ini"
    p.write_text("[game_data]\nname = x\ndate = y\ntime = 0\n\n[A]\nmass = 1\n\n[A]\nmass = 2\n")
    with pytest.raises(ValueError):
        mapfile.load_file(str(p))
    p.write_text("[game_data]\nname = x\ndate = y\ntime = 0\n\n[A]\nmass = 1\ndensity = 2\n")
    with pytest.raises(KeyError):          # neither position/velocity nor obj: the reference raises too
        mapfile.load_file(str(p))


@pytest.mark.live
def test_builtin_solar_system_map():
    """The reference's built-in map through the bulk loader (dev container only)."""
    from volatilespace_b200 import mapfile
    path = "/root/reference/Resources/BuiltinMaps/Solar System.ini"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    _check_loaded(mapfile.load_file(path), golden("mapfiles"), "solar")
