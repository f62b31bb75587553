#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 physics hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload disk1m|plummer64k] [--no-extra]

Metric (BASELINE.json): body-pair interactions/sec (FP64) for the all-pairs
gravity step (accelerations + the reference integrator v += a; P += v) on the
1M-body rotating disk, rows sharded over N GPUs with one NCCL all-gather of the
positions per step.  One "step" = one tick over all N(N-1) directed pairs.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the
library's stream with inputs resident in HBM; `e2e` is the same tick through the
host-buffer path (pinned numpy buffers, H2D + D2H inside the timed region).
`--impl reference` times the CPU oracle port (oracle/vs_oracle.c, OpenMP over all
host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# FP64-pipe instructions per directed pair: row kernel 13; symmetric kernel 8
# (16 per unordered pair: 10 shared + 3 per side)
OPS_PER_PAIR = {"rows": 13, "sym": 8}
FLOPS_PER_PAIR = 13
METRIC = "body-pair interactions/sec (FP64)"
UNIT = "pairs/s"
L2_NOTE = ("GPU arm: L2 flushed between timed steps (256 MiB memset on the timed stream, inside the timed "
           "region); within a step the 24 MiB of positions / masses are re-read from L2 by design "
           "(compute-bound kernel)")


def config_of(label, n):
    """The same `config` object in both arms (the driver compares them)."""
    return {"workload": label, "bodies": n, "pairs_per_step": n * (n - 1), "l2": L2_NOTE}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(name):
    from volatilespace_b200 import synth
    if name == "disk1m":
        s = synth.rotating_disk(1048576)
        label = "1M-body rotating disk (2-D), seed 20240602, MODE_ALLPAIRS"
    elif name == "plummer64k":
        s = synth.plummer2d(65536)
        label = "64k-body 2-D Plummer disk, seed 20240601, MODE_ALLPAIRS"
    else:
        raise SystemExit("unknown workload " + name)
    return s, label


def pinned_copy(a):
    """numpy array backed by pinned host memory (torch allocator)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    out = t.numpy()
    out[...] = a
    return out, t


# ---------------------------------------------------------------------- CPU legs
def cpu_pairs_per_s(s, budget_s, threads=None):
    """Oracle port on a bounded sample: R target rows x all N sources, all host cores.
    R is grown until one run takes about budget_s seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    L = orc.lib()
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override)
    L.vso_set_num_threads(int(threads) if threads else len(os.sched_getaffinity(0)))
    cores = L.vso_num_threads()
    n = len(s["mass"])
    rows = max(cores * 4, 32)
    while True:
        rows = min(n, max(cores, rows // cores * cores))
        t0 = time.perf_counter()
        orc.allpairs_accel(s["mass"], s["pos"], s["gc"], 0, rows)
        dt = time.perf_counter() - t0
        if dt >= 0.6 * budget_s or rows >= n:
            break
        rows = int(rows * min(8.0, 1.05 * budget_s / max(dt, 1e-6)))
    return rows * (n - 1) / dt, cores, rows, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s, label = make_workload(args.workload)
    n = len(s["mass"])
    per_step = 4.0 if args.steps + args.warmup > 4 else 8.0
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, rows, dt = cpu_pairs_per_s(s, per_step)
        if i >= args.warmup:
            vals.append((v, dt, rows))
    value = float(np.mean([v for v, _, _ in vals]))
    ms = float(np.mean([dt for _, dt, _ in vals])) * 1e3
    sample = "%d target rows x %d sources per step (oracle/vs_oracle.c, OpenMP)" % (vals[-1][2], n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(label, n),
        "note": "CPU port of the reference pair formula (gravity_one per pair); the reference itself has "
                "no all-pairs code path",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------- GPU arm
_flush_buf = None


def flush_l2(torch, stream):
    """Write a 256 MiB buffer (2x the 126 MB L2) on the timed stream: every step starts with a
    cold L2.  Costs ~40 us per step, inside the timed region."""
    global _flush_buf
    if _flush_buf is None:
        _flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(stream):
        _flush_buf.zero_()


def time_ticks(torch, stream, fn, steps, barrier):
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    flush_l2(torch, stream)     # untimed: allocates the buffer, loads torch's fill kernel
    barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        flush_l2(torch, stream)
        fn()
    ev1.record(stream)
    ev1.synchronize()
    torch.cuda.synchronize()
    barrier()
    return ev0.elapsed_time(ev1)


def time_ticks_each(torch, stream, fn, steps):
    """Sum of the per-tick durations, CUDA events around every tick; the L2 flush between ticks is
    outside the timed windows (secondary workloads whose tick takes a few ms: a 40 us flush would
    be 2 % of it)."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    flush_l2(torch, stream)
    torch.cuda.synchronize()
    for e0, e1 in evs:
        flush_l2(torch, stream)
        e0.record(stream)
        fn()
        e1.record(stream)
    evs[-1][1].synchronize()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in evs)


def extra_workloads(torch, dev, stream, peaks, fp64_ops_peak, with_cpu=True):
    """Secondary configs of BASELINE.json, device-timed, reported under `extra`."""
    from volatilespace_b200 import synth
    from volatilespace_b200._lib import KIND_VESSEL
    out = {}
    nobar = lambda: None
    # config 1: built-in Solar System map (15 bodies), MODE_REF editor frames, host wall clock
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "solar_editor.npz"))
        from volatilespace_b200.phys_editor import Physics
        res = {}
        for name in ("frame", "frame_calls"):
            ph = Physics(dev, curve_points=64)
            ph.load_system({"gc": float(g["gc"]), "rad_mult": float(g["rad_mult"]),
                            "coi_coef": float(g["coi_coef"])},
                           {"mass": g["mass"], "den": g["den"]},
                           {"pos": g["pos0"], "vel": g["vel0"]})
            ph.kepler_basic()
            fn = getattr(ph, name)
            for _ in range(50):
                fn()
            nfr = 2000 if name == "frame" else 300
            t0 = time.perf_counter()
            for _ in range(nfr):
                fn()
            res[name + "_per_s"] = nfr / (time.perf_counter() - t0)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc
        ed = orc.EditorOracle(g["mass"], g["den"], g["pos0"], g["vel0"], float(g["gc"]),
                              float(g["rad_mult"]), float(g["coi_coef"]))
        orc.lib().vso_set_num_threads(1)
        t0 = time.perf_counter()
        for _ in range(2000):
            ed.frame()
        res["cpu_port_frames_per_s_1thread"] = 2000 / (time.perf_counter() - t0)
        orc.lib().vso_set_num_threads(os.cpu_count() or 1)
        res["note"] = ("frame = fused vsb_editor_ticks driver (1 launch, the frame record written by the "
                       "kernel into mapped host memory), frame_calls = the reference's per-method calls "
                       "(gravity / body / inelastic_collision / kepler_basic: one launch each on small "
                       "systems); eager host mirrors in both")
        out["solar_system_editor"] = res
    except Exception as e:
        out["solar_system_editor"] = {"error": str(e)[:200]}
    # config 2: 64k Plummer, both all-pairs kernels
    s = synth.plummer2d(65536)
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    pairs = 65536 * 65535
    out["plummer64k"] = {}
    for mode in ("sym", "rows"):
        dev.set_gravity_mode(mode)
        for _ in range(3):
            dev.allpairs_tick_async(s["gc"])
        ms = time_ticks_each(torch, stream, lambda: dev.allpairs_tick_async(s["gc"]), 16) / 16
        out["plummer64k"][mode] = {
            "pairs_per_s": pairs / (ms * 1e-3), "ms_per_step": ms,
            "fp64_pipe_utilisation_at_1965mhz": pairs * OPS_PER_PAIR[mode] / (ms * 1e-3) /
            (148 * 64 * 1965e6),
            "flops13_frac_at_1965mhz": pairs * FLOPS_PER_PAIR / (ms * 1e-3) / (148 * 64 * 2 * 1965e6)}
    dev.set_gravity_mode("auto")
    # the slab-free symmetric kernel on the headline workload (VSB_SYM_MODE=lean; automatic above
    # 16 GB of slabs): same pairs, 48 MB of scratch instead of 8.6 GB
    try:
        d1 = synth.rotating_disk(1048576)
        dev.upload_bodies(d1["mass"], d1["pos"], d1["vel"])
        dev.set_gravity_mode("sym")
        os.environ["VSB_SYM_MODE"] = "lean"
        dev.allpairs_tick_async(d1["gc"])
        ms = time_ticks_each(torch, stream, lambda: dev.allpairs_tick_async(d1["gc"]), 3) / 3
        p1 = 1048576 * 1048575
        tr = measured_traffic("allpairs_symlean_kernel@disk1m@1gpu") or {}
        out["disk1m_slab_free_kernel"] = {
            "pairs_per_s": p1 / (ms * 1e-3), "ms_per_step": ms,
            "fp64_pipe_utilisation_at_1965mhz": p1 * 8 / (ms * 1e-3) / (148 * 64 * 1965e6),
            "traffic": tr.get("bytes"), "traffic_source": tr.get("source"),
            "note": "vsb_gravity_sym_lean.cu: persistent CTAs, ordered accumulation into acc[]; "
                    "bit-identical run to run"}
    except Exception as e:
        out["disk1m_slab_free_kernel"] = {"error": str(e)[:200]}
    finally:
        os.environ.pop("VSB_SYM_MODE", None)
        dev.set_gravity_mode("auto")
    if with_cpu:
        v, cores, rows, dtc = cpu_pairs_per_s(s, 4.0)
        out["plummer64k"]["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d target rows x 65536 sources (%.1f s), oracle/vs_oracle.c OpenMP" % (rows, dtc)}
    # MODE_REF (the reference's own parent/COI model) at the size of config 3: one
    # editor gravity tick = find_parents + gravity_one per body + chain composition + integrator
    try:
        s = synth.rotating_disk(1048576)
        r = np.hypot(s["pos"][:, 0], s["pos"][:, 1])
        coi = r * (s["mass"] / s["mass"][0]) ** 0.4      # kepler_basic's COI of a circular orbit
        coi[0] = 0.0
        dev.upload_bodies(s["mass"], s["pos"], s["vel"], coi=coi, rel_vel=s["vel"])
        order = np.argsort(s["mass"], kind="stable")[-1::-1]
        res = {"bodies": 1048576}
        for mode, reps in (("grid", 10), ("scan", 2)):
            os.environ["VSB_PARENTS"] = mode
            dev.gravity_ref(s["gc"], order)
            dev.sync()
            ms = time_ticks(torch, stream, lambda: dev.gravity_ref(s["gc"], None), reps, nobar) / reps
            res[mode] = {"ms_per_tick": ms, "body_ticks_per_s": 1048576 / (ms * 1e-3)}
        os.environ.pop("VSB_PARENTS", None)
        # A7 simplified_orbit_coi on the same state (load / add / delete / twice per merge)
        dev.simplified_orbit_coi(s["gc"], 0.4, order)
        t0 = time.perf_counter()
        dev.simplified_orbit_coi(s["gc"], 0.4, order)
        res["simplified_orbit_coi_ms"] = (time.perf_counter() - t0) * 1e3
        res["note"] = ("grid = sort-based candidate pass for find_parents (default from 4096 bodies), "
                       "scan = the exhaustive N^2/2 point-in-COI scan; identical parents")
        out["disk1m_mode_ref_tick"] = res
    except Exception as e:
        os.environ.pop("VSB_PARENTS", None)
        out["disk1m_mode_ref_tick"] = {"error": str(e)[:200]}
    # config 5: collision detection at 256k bodies, first tick of the collapsing cluster and a
    # no-hit variant (radii / 100: worst case for the exhaustive scan)
    s = synth.collapsing_cluster(262144)
    from volatilespace_b200.phys_editor import body_radius
    rad = body_radius(s["mass"], s["den"], s["rad_mult"])
    coll = {}
    for label, r in (("first_tick", rad), ("no_hit", rad * 0.01)):
        dev.upload_bodies(s["mass"], s["pos"], s["vel"], rad=r)
        for mode in ("grid", "brute"):
            os.environ["VSB_COLLISION"] = mode
            hit = dev.check_collision()
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                hit = dev.check_collision()
            dt = (time.perf_counter() - t0) / reps
            coll[label + "_" + mode] = {"ms_per_tick": dt * 1e3, "bodies_per_s": 262144 / dt,
                                        "pair": None if hit[0] is None else list(hit)}
    os.environ.pop("VSB_COLLISION", None)
    coll["timing"] = "host wall clock per vsb_check_collision call incl. result read-back"
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc            # cpu_baseline leg only
        orc.lib().vso_set_num_threads(len(os.sched_getaffinity(0)))
        cores = orc.lib().vso_num_threads()
        cb = {"cores": cores, "kind": "port", "unit": "bodies/s",
              "sample": "oracle check_collision (phys_editor.py:86-97, 547-551 restated, OpenMP over "
                        "rows with the reference's early exit), one full call per variant"}
        for label, r in (("first_tick", rad), ("no_hit", rad * 0.01)):
            t0 = time.perf_counter()
            hit = orc.check_collision(s["mass"], s["pos"], r)
            dtc = time.perf_counter() - t0
            cb[label] = {"ms_per_tick": dtc * 1e3, "bodies_per_s": 262144 / dtc,
                         "pair": None if hit[0] is None else [int(hit[0]), int(hit[1])]}
        cb["value"] = cb["no_hit"]["bodies_per_s"]
        coll["cpu_baseline"] = cb
    out["cluster256k_collision"] = coll
    # config 5 end to end: gravity(); body(); inelastic_collision() ticks with merges applied
    from volatilespace_b200.phys_editor import AllPairsPhysics
    ph = AllPairsPhysics(dev)
    ph.load_system({"gc": s["gc"], "rad_mult": s["rad_mult"]}, {"mass": s["mass"], "den": s["den"]},
                   {"pos": s["pos"], "vel": s["vel"]})
    merges = []
    ph.tick()
    t0 = time.perf_counter()
    nt = 12
    for t in range(nt):
        m = ph.tick()
        if m is not None:
            merges.append([t + 1, int(m[0]), int(m[1])])
    dt = (time.perf_counter() - t0) / nt
    out["cluster256k_ticks"] = {"ms_per_tick": dt * 1e3, "body_ticks_per_s": len(ph.mass) / dt,
                                "merges_in_%d_ticks" % nt: len(merges), "first_merges": merges[:4],
                                "note": "AllPairsPhysics.tick: all-pairs gravity + integrator, "
                                        "host np.cbrt radius, collision check, merge + compaction"}
    # config 4: 4M Kepler objects + 512-point curves (fused vessel tick), HBM roofline
    n, P = 4194304, 512
    try:
        k = synth.kepler_objects(n)
        dev.kepler_load(KIND_VESSEL, P, k["a"], k["b"], k["f"], k["ecc"], k["pea"], k["n"],
                        k["dr"], k["ma"], k["ea"], k["ref"])
        dev.kepler_tick(KIND_VESSEL, 1.0, k["body_pos"])
        for _ in range(2):
            dev.kepler_tick_async(KIND_VESSEL, 1.0)
        ms = time_ticks_each(torch, stream, lambda: dev.kepler_tick_async(KIND_VESSEL, 1.0), 5) / 5
        bytes_per_obj = 128 + 16 * P
        gbs = n * bytes_per_obj / (ms * 1e-3) / 1e9
        out["kepler4m"] = {"object_ticks_per_s": n / (ms * 1e-3), "ms_per_tick": ms,
                           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                                        "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                        "traffic": (measured_traffic("vessel_tick_kernel@kepler4m") or {}).get("bytes"),
                                        "traffic_source": (measured_traffic("vessel_tick_kernel@kepler4m") or {}).get("source"),
                                        "bytes_per_object_tick": bytes_per_obj}}
        # end to end through the host-facing calls, as game.py uses them: tick with the parent
        # positions from the host, cull on screen, gather + copy only the visible curves
        from volatilespace_b200._lib import K_POS
        sb = np.array([[-6.0e4, 4.0e4], [6.0e4, -4.0e4]])

        def game_tick():
            dev.kepler_tick(KIND_VESSEL, 1.0, k["body_pos"])
            vis = dev.kepler_culling(KIND_VESSEL, n, 10 / 0.01, sb, 0.01)
            return vis, dev.kepler_curves_gather(KIND_VESSEL, vis[:2000], P, reuse=True)
        game_tick()                      # first call: page-locks the staging buffers
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            vis, cur = game_tick()
        dt = (time.perf_counter() - t0) / reps
        out["kepler4m"]["e2e"] = {
            "object_ticks_per_s": n / dt, "ms_per_tick": dt * 1e3, "visible_objects": int(len(vis)),
            "curves_copied": int(len(cur)), "h2d_bytes_per_tick": 16 * len(k["body_pos"]) + 32,
            "d2h_bytes_per_tick": int(8 * len(vis) + 16 * P * len(cur)),
            "path": "vsb_kepler_tick + vsb_kepler_culling + vsb_kepler_curves_gather (host arrays; "
                    "results land in the Device's page-locked staging buffers)"}
        dev.kepler_free(KIND_VESSEL)
        if with_cpu:
            # CPU port of the same tick (vessel_move + curve_points + curve_move_to,
            # phys_vessel.py:477-503, phys_shared.py:202-222) on a bounded sample of the objects
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as orc            # cpu_baseline leg only
            ns = 65536
            sub = {key: np.ascontiguousarray(k[key][:ns]) for key in
                   ("a", "b", "f", "ecc", "pea", "n", "dr", "ma", "ea", "ref")}
            tt = np.linspace(-np.pi, np.pi, P)

            def cpu_tick():
                orc.vessel_move(1.0, sub["ref"], k["body_pos"], sub["a"], sub["b"], sub["f"], sub["ecc"],
                                sub["pea"], sub["n"], sub["dr"], sub["ma"], sub["ea"])
                cur = orc.curves_all(sub["ecc"], sub["a"], sub["b"], sub["pea"], tt)
                orc.curve_move_to(cur, k["body_pos"], sub["ref"], sub["f"], sub["pea"])
            cb = {"kind": "port", "unit": "object-ticks/s",
                  "sample": "%d of the 4194304 objects x %d curve points per tick (O(N): scales "
                            "linearly), oracle/vs_oracle.c" % (ns, P)}
            for label, threads in (("1_thread", 1), ("all_cores", len(os.sched_getaffinity(0)))):
                orc.lib().vso_set_num_threads(threads)
                cpu_tick()
                t0 = time.perf_counter()
                cpu_tick()
                dtc = time.perf_counter() - t0
                cb[label] = {"object_ticks_per_s": ns / dtc, "cores": orc.lib().vso_num_threads(),
                             "ms_per_sample_tick": dtc * 1e3}
            cb["value"], cb["cores"] = cb["all_cores"]["object_ticks_per_s"], cb["all_cores"]["cores"]
            out["kepler4m"]["cpu_baseline"] = cb
    except Exception as e:   # e.g. not enough free HBM on a shared device
        out["kepler4m"] = {"error": str(e)[:200]}

    # phys_vessel.Physics.points for the 160 vessels of the live-reference fixture
    # (the reference itself: ~110 s for this call sequence, dominated by four vessels whose serial
    # COI-entry march runs ~1e8 steps); CPU port timed on the short-march vessels only
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "vessel_points.npz"))
        v12, b12 = g["vessel12"], g["body12"]
        nv = len(v12)
        res = [np.full((nv, w), np.nan) for w in (2, 2, 2, 6)]
        args = (v12, g["v_ref"], b12, g["body_ref"])
        dev.host_vessel_points(*args, *res, steps_limit=300)
        t0 = time.perf_counter()
        dev.host_vessel_points(*args, *res, steps_limit=300)
        t_all = time.perf_counter() - t0
        chunks, rerun, serial, refused = dev.march_stats()
        per = np.where(v12[:, 3] < 1, v12[:, 11], np.inf)
        kids = [np.flatnonzero((g["body_ref"] == r) & (np.arange(len(b12)) != 0)) for r in g["v_ref"]]
        cost = np.array([sum(max(min(5 * p / b12[k, 9], p / 2), p / 30) for k in ks) if len(ks) else 0.0
                         for p, ks in zip(per, kids)])
        sel = np.flatnonzero(cost < 2e5)
        t0 = time.perf_counter()
        dev.host_vessel_points(*args, *res, sel=sel, steps_limit=300)
        t_sel = time.perf_counter() - t0
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc            # cpu_baseline leg only
        t0 = time.perf_counter()
        orc.vessel_points(*args, sel=sel, steps_limit=300)
        t_cpu = time.perf_counter() - t0
        out["vessel_points160"] = {
            "gpu_ms_all_vessels": t_all * 1e3, "march_chunks": chunks, "march_chunks_rerun": rerun,
            "serial_fallbacks": serial, "short_march_vessels": int(len(sel)),
            "gpu_ms_short": t_sel * 1e3, "cpu_port_ms_short": t_cpu * 1e3, "cpu_cores": 1,
            "note": "host arrays in and out (vsb_host_vessel_points); the C port needs ~100 s for "
                    "all 160 vessels on one core, the reference ~110 s (oracle/gen_golden.py "
                    "vessel_points, dev container)"}
    except Exception as e:
        out["vessel_points160"] = {"error": str(e)[:200]}

    # SURVEY 8f N2, per-tick scans: check_warp + predict_enter_coi_service + cross_coi detection on
    # one recorded state of the reference's game loop (140 vessels) and on the same state tiled to
    # 1 M vessels, through the host-array calls (uploads included)
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "vessel_events.npz"))
        v, p = g["cw0_v"], g["cw0_p"]
        res = {}
        for label, reps_tile in (("140", 1), ("1m", 7490)):
            vv, pp = np.tile(v, (reps_tile, 1)), np.tile(p, (reps_tile, 1))
            cols = [np.ascontiguousarray(vv[:, c]) for c in range(14)]
            bi, ba, cl, ce = (np.ascontiguousarray(x) for x in (pp[:, 0:2], pp[:, 2:4], pp[:, 4:6], pp[:, 6:12]))
            hold = np.zeros(len(vv), dtype=np.uint8)
            def scans():
                dev.host_check_warp(cols[3], cols[6], cols[8], cols[7], cols[5], bi, ba, cl, ce, 5.0, hold)
                dev.host_enter_coi_service(cols[6], cols[9], cols[8], ce, False)
                dev.host_cross_scan(cols[3], cols[6], cols[9], cols[8], bi, cl, ce)
            scans()
            t0 = time.perf_counter()
            for _ in range(5):
                scans()
            res["ms_per_tick_" + label] = (time.perf_counter() - t0) / 5 * 1e3
            res["vessels_" + label] = int(len(vv))
        # the same three scans on device-resident state (1 M vessels): nothing per-vessel over PCIe
        n1 = len(vv)
        dev.kepler_load(KIND_VESSEL, 8, cols[0], cols[1], cols[2], cols[3], cols[4], cols[5],
                        cols[6], cols[7], cols[8], np.zeros(n1, dtype=np.int64))
        dev.vessel_events_load(cols[11], cols[12], cols[13], bi, ba, cl, ce, None, cols[9])
        def resident():
            dev.vessel_check_warp(5.0)
            dev.vessel_enter_coi_service(n1, False)
            dev.vessel_cross_scan()
        resident()
        t0 = time.perf_counter()
        for _ in range(10):
            resident()
        res["ms_per_tick_1m_resident"] = (time.perf_counter() - t0) / 10 * 1e3
        # curve_segments (light / dark parts of the visible orbit curves) on the same resident state:
        # a fresh 1 M-vessel state with P = 64 curves, 16 384 visible orbits
        P = 64
        dev.kepler_load(KIND_VESSEL, P, cols[0], cols[1], cols[2], cols[3], cols[4], cols[5],
                        cols[6], cols[7], cols[8], np.zeros(n1, dtype=np.int64))
        dev.vessel_events_load(cols[11], cols[12], cols[13], bi, ba, cl, ce, None, cols[9])
        dev.kepler_tick(KIND_VESSEL, 0.0, np.zeros((1, 2)))
        vis = np.arange(0, n1, n1 // 16384, dtype=np.int64)[:16384]
        seg = dev.vessel_curve_segments(vis, P, reuse=True)
        t0 = time.perf_counter()
        for _ in range(10):
            seg = dev.vessel_curve_segments(vis, P, reuse=True)
        res["curve_segments_ms_16k_visible_p64"] = (time.perf_counter() - t0) / 10 * 1e3
        res["curve_segments_types"] = np.bincount(seg[3], minlength=4).tolist()
        dev.kepler_free(KIND_VESSEL)
        res["note"] = ("three host-array calls per tick, uploads of the per-vessel arrays included "
                       "(320 MB per tick at 1 M vessels: PCIe-bound); the reference's whole game tick "
                       "ran at ~67 ms at 140 vessels while the fixture was generated")
        out["vessel_event_scans"] = res
    except Exception as e:
        out["vessel_event_scans"] = {"error": str(e)[:200]}
    return out


def multi_gpu_extras(torch, dist, dev, stream, rank, world, barrier):
    """The other two rows of SURVEY 8e at N > 1 (every rank takes part; max over ranks):
    Kepler propagation in object blocks (no collective) and collision detection at 256k bodies
    (row-split exhaustive scan + 64-bit MIN all-reduce, and the redundant candidate pass)."""
    from volatilespace_b200 import synth
    from volatilespace_b200._lib import KIND_VESSEL
    from volatilespace_b200.dist import ShardedAllPairsPhysics, partition
    from volatilespace_b200.phys_editor import body_radius
    out = {}

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, P = 4194304, 512
    rpr, ranges = partition(n, world)
    mine = ranges[rank][1] - ranges[rank][0]
    k = synth.kepler_objects(mine, seed=20240603 + rank)      # this rank's block of the 4M objects
    dev.kepler_load(KIND_VESSEL, P, k["a"], k["b"], k["f"], k["ecc"], k["pea"], k["n"], k["dr"],
                    k["ma"], k["ea"], k["ref"])
    dev.kepler_tick(KIND_VESSEL, 1.0, k["body_pos"])
    for _ in range(2):
        dev.kepler_tick_async(KIND_VESSEL, 1.0)
    ms = time_ticks(torch, stream, lambda: dev.kepler_tick_async(KIND_VESSEL, 1.0), 5, barrier) / 5
    ms = max_over_ranks(ms)
    out["kepler4m_sharded"] = {"ms_per_tick": ms, "object_ticks_per_s": n / (ms * 1e-3),
                               "objects_per_rank": rpr, "collective": "none"}
    dev.kepler_free(KIND_VESSEL)

    s = synth.collapsing_cluster(262144)
    rad = body_radius(s["mass"], s["den"], s["rad_mult"]) * 0.01           # no hit: full scan
    coll = {}
    for mode in ("rows", "grid"):
        ph = ShardedAllPairsPhysics(dev, rank, world, collisions=mode)
        ph.load_system({"gc": s["gc"], "rad_mult": s["rad_mult"]}, {"mass": s["mass"], "den": s["den"]},
                       {"pos": s["pos"], "vel": s["vel"]})
        ph.dev.set(_f_rad(), rad)
        assert ph.check_collision() == (None, None)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            ph.check_collision()
        coll[mode + "_ms_per_check"] = max_over_ranks((time.perf_counter() - t0) / 5 * 1e3)
    coll["note"] = ("262144 bodies, no hit (worst case), host-visible result per check; rows = "
                    "exhaustive scan split by row blocks + MIN all-reduce of one 64-bit key, grid = "
                    "sort-based candidate pass on every rank")
    out["collision256k_sharded"] = coll
    return out


def _f_rad():
    from volatilespace_b200._lib import F_RAD
    return F_RAD


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel on a named workload, from
    the committed ncu captures (profiles/traffic.json: {key: {"bytes": ..., "source": file}})."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except OSError:
        return None


def roofline(pairs_per_s, fp64_flops_microbench, clocks, ms_accel, ops_per_pair, kernel, traffic_key):
    """FP64 roofline of the gravity kernel.  `achieved` / `frac` use the ALGORITHMIC count of
    SURVEY 8d (13 flop per directed pair); the share of FP64-pipe issue slots the kernel fills
    (what ncu's sm__inst_executed_pipe_fp64 reports, the north star's own definition) is the
    separate key `fp64_pipe_utilisation`.  MEASURED_PEAKS.json has no FP64 figure, so the peak is
    the pipe's issue peak: 64 FP64 lanes/SM/clk x 148 SMs x the SM clock observed during the
    timed region x 2 flop (a DFMA per lane per clock)."""
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    lanes_per_s = 148 * 64 * sm_clock * 1e6
    peak = 2 * lanes_per_s
    tr = measured_traffic(traffic_key)
    return {
        "bound": "fp64", "achieved": pairs_per_s * FLOPS_PER_PAIR / 1e12, "peak": peak / 1e12,
        "unit": "TFLOP/s", "frac": pairs_per_s * FLOPS_PER_PAIR / peak,
        "flops_per_pair": FLOPS_PER_PAIR,
        "fp64_pipe_utilisation": pairs_per_s * ops_per_pair / lanes_per_s,
        "fp64_pipe_instr_per_pair": ops_per_pair,
        "traffic": None if tr is None else tr["bytes"],
        "traffic_source": None if tr is None else tr["source"],
        "kernel": kernel, "ms_per_launch": ms_accel,
        "peak_source": "64 FP64 lanes/SM/clk x 148 SMs x %.0f MHz observed x 2 (issue peak of the "
                       "FP64 pipe; not in MEASURED_PEAKS.json)" % sm_clock,
        "dfma_microbench_tflops": fp64_flops_microbench / 1e12,
    }


def parity_check(dev, s, rows=64, acc=None):
    """Accelerations of `rows` sampled target rows of the state now on the device against the
    CPU oracle (each row: N reference pair evaluations).  Run AFTER the timed region."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc            # checker only
    from volatilespace_b200._lib import F_ACC, F_POS
    n = len(s["mass"])
    pos = dev.get(F_POS)
    if acc is None:
        acc = dev.get(F_ACC)
    pick = np.unique(np.concatenate(([1, n - 1], np.random.default_rng(11).choice(n, rows - 2, replace=False))))
    worst = 0.0
    for r in pick:
        want = orc.allpairs_accel(s["mass"], pos, s["gc"], int(r), int(r) + 1)[0]
        worst = max(worst, float(np.linalg.norm(acc[r] - want) / np.linalg.norm(want)))
    return {"rows": int(len(pick)), "max_rel": worst, "tol": 1e-11, "ok": bool(worst < 1e-11),
            "what": "acceleration of sampled target rows after the timed ticks vs oracle/vs_oracle.c"}


def run_b200(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        barrier = lambda: dist.barrier()
    else:
        dist = None
        barrier = lambda: None

    from volatilespace_b200 import _lib
    from volatilespace_b200._lib import F_ACC
    from volatilespace_b200.device import Device
    from volatilespace_b200.dist import ShardedAllPairs, ShardedAllPairsPhysics, ShardedAllPairsSym
    dev = Device(local)
    stream = torch.cuda.ExternalStream(dev.stream_handle(), device="cuda:%d" % local)
    peaks, peaks_src = load_peaks()
    fp64_flops_peak, fp64_ops_peak = dev.fp64_peak()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    s, label = make_workload(args.workload)
    n = len(s["mass"])
    pairs = n * (n - 1)
    if args.gravity == "sym":
        sim = ShardedAllPairsSym(dev, s["mass"], s["pos"], s["vel"], s["gc"], rank, world)
    else:
        sim = ShardedAllPairs(dev, s["mass"], s["pos"], s["vel"], s["gc"], rank, world)
    ops_per_pair = OPS_PER_PAIR[args.gravity]
    for _ in range(args.warmup):
        sim.tick()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = dev.kernel_launches()
    ms_total = time_ticks(torch, stream, sim.tick, args.steps, barrier)
    launches = dev.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms_total) / args.steps
    value = pairs / (ms_step * 1e-3)

    # kernel-only timing of the dominant kernel (accel) for the roofline, CUDA events on the
    # library's stream; max over ranks (each rank runs its share of the pairs)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = max(2, min(args.steps, 4))
    torch.cuda.synchronize()
    barrier()
    ev[0].record(stream)
    for _ in range(reps):
        if args.gravity == "sym" and world > 1:
            dev.allpairs_sym_partial_async(s["gc"], rank, world)
        else:
            _lib.check(dev.L.vsb_allpairs_accel_async(dev.h, float(s["gc"])))
    ev[1].record(stream)
    ev[1].synchronize()
    ms_accel = max_over_ranks(ev[0].elapsed_time(ev[1]) / reps)

    # parity, after the timed region (the bench carries its own proof): the state the
    # timed ticks left on the device, 64 sampled rows against the oracle; at N > 1 additionally
    # the all-reduced accelerations against the one-GPU kernel run on rank 0's replica
    parity = None
    if not args.no_parity:
        if args.gravity == "sym" and world > 1:
            dev.allpairs_sym_partial_async(s["gc"], rank, world)
            with torch.cuda.stream(sim._stream):
                dist.all_reduce(sim._acc_t, group=None)
            sim._stream.synchronize()
            acc_multi = dev.get(F_ACC)
        else:
            if world > 1:
                dev.set_rows(0, n)          # every row on this rank, for the check only
            dev.allpairs_accel(s["gc"])
            acc_multi = dev.get(F_ACC)
            if world > 1:
                dev.set_rows(sim.row0, sim.row1)
        if rank == 0:
            parity = parity_check(dev, s, 64, acc_multi)
            if world > 1 and args.gravity == "sym":
                dev.allpairs_accel(s["gc"])                 # one-GPU kernel, same positions
                one = dev.get(F_ACC)
                rel = np.linalg.norm(acc_multi - one, axis=1) / np.maximum(np.linalg.norm(one, axis=1), 1e-300)
                parity["vs_one_gpu_kernel_max_rel"] = float(rel.max())
                parity["ok"] = bool(parity["ok"] and rel.max() < 1e-11)
        barrier()

    # end to end, tick_host: pinned numpy buffers, each rank moves ITS row slice over PCIe
    hp, _k1 = pinned_copy(s["pos"])
    hv, _k2 = pinned_copy(s["vel"])
    hm, _k3 = pinned_copy(s["mass"])
    e2e_steps = max(1, min(args.steps, 3))

    def wall(fn):
        fn()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    dt_host = wall(lambda: sim.tick_host(hm, hp, hv))
    h2d, d2h = sim.host_bytes_per_tick()
    e2e_tick_host = {"value": pairs / dt_host, "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": dt_host * 1e3,
                     "path": "%s.tick_host -> vsb_bodies_set / all-pairs tick / vsb_bodies_get "
                             "(pinned host buffers, per-rank row slices)" % type(sim).__name__}
    # end to end through the DROP-IN class (phys_editor.AllPairsPhysics; ShardedAllPairsPhysics at
    # N > 1): per step the caller's mass / pos / vel go host -> device from the class's page-locked
    # mirrors, gravity() runs the tick, pos / vel come back into the mirrors.  The class keeps FULL
    # host mirrors on every rank, so at N > 1 every rank moves all bodies over PCIe; there the
    # headline `e2e` is tick_host (each rank moves its slice of the rows), the class path is
    # reported beside it as `e2e_dropin`.
    e2e, e2e_dropin = e2e_tick_host, None
    if args.gravity == "sym":
        ph = ShardedAllPairsPhysics(dev, rank, world)
        ph.load_system({"gc": s["gc"], "rad_mult": s["rad_mult"]}, {"mass": s["mass"], "den": s["den"]},
                       {"pos": s["pos"], "vel": s["vel"]})

        def dropin_step():
            ph.push("mass", "pos", "vel")
            ph.gravity()
            ph.inner.pull("pos", "vel")
        dt_cls = wall(dropin_step)
        e2e_dropin = {"value": pairs / dt_cls, "unit": UNIT, "h2d_bytes_per_step": 40 * ((n + world - 1) // world),
                      "d2h_bytes_per_step": 32 * n, "ms_per_step": dt_cls * 1e3,
                      "path": "phys_editor.AllPairsPhysics%s: mirrors -> device (mass, pos, vel), gravity(), "
                              "pull(pos, vel); page-locked mirrors; at N > 1 each rank uploads its block of rows (NCCL "
                              "all-gather completes the replicas) and downloads all bodies" %
                              ("" if world == 1 else " via dist.ShardedAllPairsPhysics")}
        if world == 1:
            e2e = e2e_dropin
        dev.set_gravity_mode("sym")

    extra_multi = None
    if world > 1 and not args.no_extra:
        extra_multi = multi_gpu_extras(torch, dist, dev, stream, rank, world, barrier)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    kernel_name = "allpairs_sym_kernel" if args.gravity == "sym" else "allpairs_accel_kernel"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(label, n),
        "kernel": "symmetric (unordered pairs, 8 FP64 instr per directed pair)"
        if args.gravity == "sym" else "rows (13 FP64 instr per directed pair)",
        "parallelism": ("1 GPU" if world == 1 else
                        "unordered block pairs/%d + NCCL all-reduce(acc)" % world
                        if args.gravity == "sym" else "rows/%d + NCCL all-gather(pos)" % world),
        "steps_per_s": 1e3 / ms_step,
        "clocks": clocks,
        "e2e": e2e,
        "e2e_tick_host": e2e_tick_host,
        "e2e_dropin": e2e_dropin,
        "gpu_launches": int(launches),
        "parity": parity,
        "sym_sched": _lib.sym_schedule(),
        # per GPU: each rank's share of the pairs over the slowest rank's kernel time, against ONE
        # GPU's peak
        "roofline": dict(roofline(pairs / world / (ms_accel * 1e-3), fp64_flops_peak, clocks, ms_accel,
                                  ops_per_pair, kernel_name,
                                  "%s@%s@%dgpu" % (kernel_name, args.workload, world)),
                         scope="per GPU (1/%d of the pairs per launch)" % world),
        "peaks_source": peaks_src,
    }
    if line["sym_sched"] != "resched":
        sys.stderr.write("bench.py: WARNING: symmetric kernel runs ptxas's schedule: %s\n" % line["sym_sched"])
    if not args.no_extra and world == 1:
        line["extra"] = extra_workloads(torch, dev, stream, peaks, fp64_ops_peak, not args.no_cpu)
    if extra_multi is not None:
        line["extra"] = extra_multi
    if not args.no_cpu and world == 1:
        v, cores, rows, dtc = cpu_pairs_per_s(s, 12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d target rows x %d sources (%.1f s), oracle/vs_oracle.c "
                                          "OpenMP" % (rows, n, dtc)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="disk1m", choices=["disk1m", "plummer64k"])
    ap.add_argument("--gravity", default="sym", choices=["sym", "rows"],
                    help="all-pairs kernel: symmetric (default) or the shard-invariant row kernel")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
