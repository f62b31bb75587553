"""Small fixed workloads for ncu captures (see profiles/README.md).

    python tools/profile_run.py gravity64k | gravity64k_sym | gravity256k | gravity1m | gravity1m_lean | kepler4m | moderef1m | soc1m | convert1m | kepler1m | collision256k | parents64k | points160
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import KIND_VESSEL  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402


def main():
    mode = sys.argv[1]
    dev = Device(0)
    t0 = time.perf_counter()
    if mode == "gravity64k":
        s = synth.plummer2d(65536)
        dev.upload_bodies(s["mass"], s["pos"], s["vel"])
        dev.allpairs_step(s["gc"], 6)
    elif mode == "gravity256k":
        s = synth.rotating_disk(262144)
        dev.upload_bodies(s["mass"], s["pos"], s["vel"])
        dev.allpairs_step(s["gc"], 3)
    elif mode == "moderef1m":
        s = synth.rotating_disk(1 << 20)
        r = np.hypot(s["pos"][:, 0], s["pos"][:, 1])
        coi = r * (s["mass"] / s["mass"][0]) ** 0.4
        coi[0] = 0.0
        dev.upload_bodies(s["mass"], s["pos"], s["vel"], coi=coi, rel_vel=s["vel"])
        order = np.argsort(s["mass"], kind="stable")[-1::-1]
        dev.gravity_ref(s["gc"], order)
        for _ in range(3):
            dev.gravity_ref(s["gc"], None)
    elif mode == "soc1m":
        s = synth.rotating_disk(1 << 20)
        dev.upload_bodies(s["mass"], s["pos"], s["vel"])
        order = np.argsort(s["mass"], kind="stable")[-1::-1]
        dev.simplified_orbit_coi(s["gc"], 0.4, order)
        t1 = time.perf_counter()
        dev.simplified_orbit_coi(s["gc"], 0.4, order)
        print("simplified_orbit_coi at 1M: %.1f ms" % ((time.perf_counter() - t1) * 1e3))
    elif mode == "convert1m":
        from volatilespace_b200 import convert
        s = synth.rotating_disk(1 << 20)
        for rep in range(2):
            t1 = time.perf_counter()
            k = convert.to_kepler(s["mass"], {"pos": s["pos"], "vel": s["vel"]}, s["gc"], 0.4, device=dev)
            t2 = time.perf_counter()
            print("to_kepler at 1M: %.1f ms (host arrays in and out)" % ((t2 - t1) * 1e3))
        print("refs != 0:", int((k["ref"] != 0).sum()))
    elif mode == "kepler1m":
        k = synth.kepler_objects(1 << 20)
        dev.kepler_load(KIND_VESSEL, 512, k["a"], k["b"], k["f"], k["ecc"], k["pea"], k["n"],
                        k["dr"], k["ma"], k["ea"], k["ref"])
        for _ in range(5):
            dev.kepler_tick(KIND_VESSEL, 1.0, k["body_pos"])
    elif mode == "kepler4m":
        k = synth.kepler_objects(1 << 22)
        dev.kepler_load(KIND_VESSEL, 512, k["a"], k["b"], k["f"], k["ecc"], k["pea"], k["n"],
                        k["dr"], k["ma"], k["ea"], k["ref"])
        for _ in range(4):
            dev.kepler_tick(KIND_VESSEL, 1.0, k["body_pos"])
    elif mode in ("gravity1m", "gravity1m_lean", "gravity64k_sym"):
        if mode == "gravity1m_lean":
            os.environ["VSB_SYM_MODE"] = "lean"
        s = synth.plummer2d(65536) if mode == "gravity64k_sym" else synth.rotating_disk(1 << 20)
        dev.upload_bodies(s["mass"], s["pos"], s["vel"])
        dev.set_gravity_mode("sym")
        dev.allpairs_step(s["gc"], 3)
    elif mode == "gravity1m_rank0of8":
        s = synth.rotating_disk(1 << 20)
        dev.upload_bodies(s["mass"], s["pos"], s["vel"])
        dev.set_gravity_mode("sym")
        for _ in range(3):
            dev.allpairs_sym_partial_async(s["gc"], 0, 8)
    elif mode == "collision256k":
        from volatilespace_b200.phys_editor import body_radius
        s = synth.collapsing_cluster(262144)
        rad = body_radius(s["mass"], s["den"], s["rad_mult"]) * 0.01    # no hit: full scan
        dev.upload_bodies(s["mass"], s["pos"], s["vel"], rad=rad)
        for _ in range(4):
            print(dev.check_collision())
    elif mode == "parents64k":
        s = synth.plummer2d(65536)
        coi = np.where(np.arange(65536) % 7 == 0, 3.0e3, 0.0)
        dev.upload_bodies(s["mass"], s["pos"], s["vel"], coi=coi)
        order = np.argsort(s["mass"])[-1::-1]
        for _ in range(4):
            dev.find_parents(order, want=False)
    elif mode == "points160":
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                                 "vessel_points.npz"))
        nv = len(g["vessel12"])
        res = [np.full((nv, w), np.nan) for w in (2, 2, 2, 6)]
        dev.host_vessel_points(g["vessel12"], g["v_ref"], g["body12"], g["body_ref"], *res)
        print("march stats", dev.march_stats())
    else:
        raise SystemExit("unknown mode")
    dev.sync()
    print(mode, "done in %.3f s, launches %d" % (time.perf_counter() - t0, dev.kernel_launches()))


if __name__ == "__main__":
    main()
