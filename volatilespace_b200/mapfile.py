"""Bulk map / savegame I/O in the reference's INI format (second half):
`load_file` and `save_file` with the signatures and return layout of
volatilespace/peripherals.py:237-359 / :362-482, in O(N).

The reference grows every array with `np.append` / `np.vstack` once per body (O(N^2) copies,
measured in the dev container: load 1.8 s at 10 000 bodies, 24 s at 40 000, i.e. hours at 10^6;
this module: 0.17 s and 1.0 s) and tests every new section name against the list of all
sections; here the file is scanned once into Python lists and converted at the end,
and section names are tracked in a set.  Parsing rules follow configparser as the reference uses
it (keys are case-insensitive, `key = value` or `key : value`, `#` / `;` comment lines, values
stripped), without its interpolation.
"""
import numpy as np

SIM_CONFIG_DEFAULTS = {          # volatilespace/defaults.py:58-65
    "gc": 1.0,
    "rad_mult": 179.0,
    "mass_thermal_mult": 2 * 10**26,
    "coi_coef": 0.7,
    "vessel_scale": 0.1,
    "min_planet_mass": 200,
}

_RESERVED = ("game_data", "config")


def _sections(path):
    """[(section name, {key: value})] in file order."""
    out, names, cur = [], set(), None
    with open(path, "r") as f:
        for raw in f:
            line = raw.strip()
            if not line or line[0] in "#;":
                continue
            if line[0] == "[" and line[-1] == "]":
                name = line[1:-1]
                if name in names:
                    raise ValueError("section %r appears twice in %s" % (name, path))
                names.add(name)
                cur = {}
                out.append((name, cur))
                continue
            if cur is None:
                raise ValueError("%s: text before the first section" % path)
            cut = min((i for i in (line.find("="), line.find(":")) if i >= 0), default=-1)
            if cut < 0:
                raise ValueError("%s: cannot parse line %r" % (path, line))
            cur[line[:cut].strip().lower()] = line[cut + 1:].strip()
    return out


def _pair(text, conv=float):
    return [conv(x) for x in text.strip("][").split(", ")]


def _atmosphere(sec):
    try:
        return float(sec["atm_pres0"]), float(sec["atm_scale_h"]), float(sec["atm_den0"])
    except (KeyError, ValueError):
        return 0.0, 0.0, 0.0


def load_file(path):
    """-> (game_data, config, body_data, body_orb_data, vessel_data, vessel_orb_data), the tuple
    of peripherals.load_file.  A file is Newtonian (position / velocity per body) until the first
    body without them; from there on sections are Keplerian bodies or vessels (`obj = ...`)."""
    secs = _sections(path)
    by_name = dict(secs)
    gd = by_name["game_data"]
    try:
        vessel = int(gd["vessel"])
    except (KeyError, ValueError):
        vessel = None
    game_data = {"name": gd["name"].strip('"'), "date": gd["date"].strip('"'),
                 "time": float(gd["time"]), "vessel": vessel}
    try:
        config = {k: float(by_name["config"][k]) for k in SIM_CONFIG_DEFAULTS}
    except (KeyError, ValueError):
        config = dict(SIM_CONFIG_DEFAULTS)

    names, mass, den, color, atm = [], [], [], [], []
    pos, vel = [], []
    k_a, k_ecc, k_pea, k_ma, k_ref, k_dir = [], [], [], [], [], []
    v_a, v_ecc, v_pea, v_ma, v_ref, v_dir = [], [], [], [], [], []
    v_name, v_mass, v_rot_angle, v_rot_acc, v_sprite = [], [], [], [], []
    kepler = False
    for name, sec in secs:
        if name in _RESERVED:
            continue
        if not kepler:
            if "position" in sec and "velocity" in sec:
                pos.append(_pair(sec["position"]))
                vel.append(_pair(sec["velocity"]))
                names.append(name)
                mass.append(float(sec["mass"]))
                den.append(float(sec["density"]))
                color.append(_pair(sec["color"], int))
                atm.append(_atmosphere(sec))
                continue
            kepler = True
        if sec["obj"] == "body":
            names.append(name)
            mass.append(float(sec["mass"]))
            den.append(float(sec["density"]))
            color.append(_pair(sec["color"], int))
            atm.append(_atmosphere(sec))
            k_a.append(float(sec["sma"]))
            k_ecc.append(float(sec["ecc"]))
            k_pea.append(float(sec["lpe"]))
            k_ma.append(float(sec["mna"]))
            k_ref.append(int(sec["ref"]))
            k_dir.append(float(sec["dir"]))
        elif sec["obj"] == "vessel":
            v_a.append(float(sec["sma"]))
            v_ecc.append(float(sec["ecc"]))
            v_pea.append(float(sec["lpe"]))
            v_ma.append(float(sec["mna"]))
            v_ref.append(int(sec["ref"]))
            v_dir.append(float(sec["dir"]))
            v_name.append(name)
            v_mass.append(float(sec["mass"]))
            v_rot_angle.append(float(sec["rot_angle"]))
            v_rot_acc.append(float(sec["rot_acc"]))
            v_sprite.append(sec["sprite"])

    f64 = lambda x: np.array(x, dtype=np.float64)
    atm = np.array(atm, dtype=np.float64).reshape(-1, 3)
    body_data = {"name": np.array(names), "mass": f64(mass), "den": f64(den),
                 "color": np.array(color, dtype=int).reshape(-1, 3), "atm_pres0": atm[:, 0].copy(),
                 "atm_scale_h": atm[:, 1].copy(), "atm_den0": atm[:, 2].copy()}
    if kepler:
        body_orb_data = {"kepler": True, "a": f64(k_a), "ecc": f64(k_ecc), "pe_arg": f64(k_pea),
                         "ma": f64(k_ma), "ref": np.array(k_ref, dtype=int), "dir": f64(k_dir)}
    else:
        body_orb_data = {"kepler": False, "pos": f64(pos).reshape(-1, 2),
                         "vel": f64(vel).reshape(-1, 2)}
    va, ve = f64(v_a), f64(v_ecc)
    va = np.where(ve > 1, -np.abs(va), va)          # hyperbolic vessels carry a negative sma (:330)
    vessel_data = {"name": np.array(v_name), "mass": f64(v_mass), "rot_angle": f64(v_rot_angle),
                   "rot_acc": f64(v_rot_acc), "sprite": np.array(v_sprite)}
    vessel_orb_data = {"a": va, "ecc": ve, "pe_arg": f64(v_pea), "ma": f64(v_ma),
                       "ref": np.array(v_ref, dtype=int), "dir": f64(v_dir)}
    return game_data, config, body_data, body_orb_data, vessel_data, vessel_orb_data


def _unique(name, taken):
    """The reference's rule for a section name that already exists: `name 1`, `name 2`, ..."""
    cand, num = str(name), 1
    while cand in taken:
        cand = "%s %d" % (name, num)
        num += 1
    taken.add(cand)
    return cand


def save_file(path, game_data, conf, body_data, body_orb_data, vessel_data=None,
              vessel_orb_data=None):
    """Write a map / savegame as peripherals.save_file does (same sections, keys, number
    formatting and Newtonian / Keplerian choice).  Unlike the reference it does not create the
    Maps/ and Saves/ directories of the game's working directory."""
    name = game_data["name"]
    if name == "":
        name = "Unnamed"
    if name is None:                 # "keep the old name" when overwriting (:378-385)
        try:
            name = dict(_sections(path))["game_data"]["name"].strip('"')
        except (OSError, KeyError, ValueError):
            name = "New map"
    lines = ["[game_data]", "name = %s" % name, "date = %s" % game_data["date"],
             "time = %s" % game_data["time"], "vessel = %s" % game_data["vessel"], "", "[config]"]
    for key in conf:
        text = str(conf[key])
        lines.append("%s = %s" % (key, text if len(text) < 10 else "{:.5e}".format(conf[key])))
    lines.append("")
    taken = set(_RESERVED)
    kepler = not ("pos" in body_orb_data and "vel" in body_orb_data)
    mass, den, color = body_data["mass"], body_data["den"], body_data["color"]
    ap, ah, ad = body_data["atm_pres0"], body_data["atm_scale_h"], body_data["atm_den0"]
    for i, body_name in enumerate(body_data["name"]):
        lines.append("[%s]" % _unique(body_name, taken))
        if kepler:
            lines.append("obj = body")
        lines.append("mass = %s" % mass[i])
        lines.append("density = %s" % den[i])
        lines.append("color = [%s, %s, %s]" % (color[i, 0], color[i, 1], color[i, 2]))
        if ap[i] != 0 and ah[i] != 0 and ad[i] != 0:
            lines.append("atm_pres0 = %s" % ap[i])
            lines.append("atm_scale_h = %s" % ah[i])
            lines.append("atm_den0 = %s" % ad[i])
        if not kepler:
            p, v = body_orb_data["pos"][i], body_orb_data["vel"][i]
            lines.append("position = [%s, %s]" % (p[0], p[1]))
            lines.append("velocity = [%s, %s]" % (v[0], v[1]))
        else:
            o = body_orb_data
            lines += ["sma = %s" % o["a"][i], "ecc = %s" % o["ecc"][i], "lpe = %s" % o["pe_arg"][i],
                      "mna = %s" % o["ma"][i], "ref = %d" % int(o["ref"][i]),
                      "dir = %d" % int(o["dir"][i])]
        lines.append("")
    if vessel_data:
        o, d = vessel_orb_data, vessel_data
        for i, v_name in enumerate(d["name"]):
            lines.append("[%s]" % _unique(v_name, taken))
            lines += ["obj = vessel", "sma = %s" % o["a"][i], "ecc = %s" % o["ecc"][i],
                      "lpe = %s" % o["pe_arg"][i], "mna = %s" % o["ma"][i],
                      "ref = %d" % int(o["ref"][i]), "dir = %d" % int(o["dir"][i]),
                      "mass = %s" % d["mass"][i], "rot_angle = %s" % d["rot_angle"][i],
                      "rot_acc = %s" % d["rot_acc"][i], "sprite = %s" % d["sprite"][i], ""]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
