"""Multi-GPU parity check (run under torchrun on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/multigpu_check.py [n_bodies] [ticks]

Every rank runs the row-sharded all-pairs ticks (NCCL all-gather of positions) and then the
same ticks unsharded on its own GPU; gathered positions and own-row velocities must be
BIT-IDENTICAL to the single-GPU result (the summation order does not depend on the sharding).
Then: the symmetric driver (all-reduce of partial accelerations), the body-sim tick with
collisions and merges (row-split scan + 64-bit MIN all-reduce, and the redundant candidate pass)
against the single-GPU class, and Kepler propagation in object blocks (no collective).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volatilespace_b200 import synth  # noqa: E402
from volatilespace_b200._lib import F_POS, F_VEL  # noqa: E402
from volatilespace_b200.device import Device  # noqa: E402
from volatilespace_b200.dist import ShardedAllPairs  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50001
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = synth.plummer2d(n, seed=11)
    dev = Device(local)
    dev.set_gravity_mode("rows")      # the row kernel is the shard-invariant one
    sim = ShardedAllPairs(dev, s["mass"], s["pos"], s["vel"], s["gc"], rank, world)
    for _ in range(ticks):
        sim.tick()
    if world > 1:
        sim._stream.synchronize()
    pos_sharded = sim.positions()
    vel_own = sim.velocities_own()
    # unsharded on the same GPU
    dev.upload_bodies(s["mass"], s["pos"], s["vel"])
    dev.allpairs_step(s["gc"], ticks)
    pos_single = dev.get(F_POS)
    vel_single = dev.get(F_VEL)
    ok = np.array_equal(pos_sharded, pos_single) and np.array_equal(
        vel_own, vel_single[sim.row0:sim.row1])
    # end-to-end host path once
    hp, hv = s["pos"].copy(), s["vel"].copy()
    sim2 = ShardedAllPairs(dev, s["mass"], s["pos"], s["vel"], s["gc"], rank, world)
    for _ in range(ticks):
        sim2.tick_host(s["mass"], hp, hv)
    ok_host = np.array_equal(hp[sim2.row0:sim2.row1], pos_single[sim2.row0:sim2.row1]) and \
        np.array_equal(hv[sim2.row0:sim2.row1], vel_single[sim2.row0:sim2.row1])
    # the symmetric driver's host path: own slices only
    from volatilespace_b200.dist import ShardedAllPairsSym
    hp2, hv2 = s["pos"].copy(), s["vel"].copy()
    sym_h = ShardedAllPairsSym(dev, s["mass"], s["pos"], s["vel"], s["gc"], rank, world)
    for _ in range(ticks):
        sym_h.tick_host(s["mass"], hp2, hv2)
    ok_host = ok_host and np.allclose(hp2[sym_h.row0:sym_h.row1],
                                      pos_single[sym_h.row0:sym_h.row1], rtol=1e-11, atol=1e-9)
    dev.set_gravity_mode("rows")
    # symmetric kernel over the ranks (all-reduce of partial accelerations): tolerance parity
    # with the row kernel, and identical replicas on every rank
    from volatilespace_b200.dist import ShardedAllPairsSym
    sym = ShardedAllPairsSym(dev, s["mass"], s["pos"], s["vel"], s["gc"], rank, world)
    for _ in range(ticks):
        sym.tick()
    if world > 1:
        sym._stream.synchronize()
    pos_sym = sym.positions()
    ok_sym = np.allclose(pos_sym, pos_single, rtol=1e-11, atol=1e-9)
    t = torch.from_numpy(pos_sym).cuda()
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    ok_rep = bool(torch.equal(t, t0))
    # body-sim tick WITH collisions over the ranks: merge list and final state of a
    # collapsing cluster against the single-GPU class, for both detection modes
    from volatilespace_b200.dist import ShardedAllPairsPhysics, ShardedKepler
    from volatilespace_b200.phys_editor import AllPairsPhysics
    c = synth.collapsing_cluster(20000, seed=5, radius=2.5e4)
    conf = {"gc": c["gc"], "rad_mult": c["rad_mult"]}
    bd = {"mass": c["mass"], "den": c["den"]}
    bo = {"pos": c["pos"], "vel": c["vel"]}
    dev.set_gravity_mode("sym")
    single = AllPairsPhysics(dev)
    single.load_system(conf, bd, bo)
    # dense on purpose (a merge every tick, more overlaps queued): the merge list must be identical;
    # positions are compared after 5 ticks -- later a pair of overlapping light bodies passes within
    # 1e-3 of each other and amplifies the last-bit difference of the two summation orders
    def run(obj):
        merges, pos5 = [], None
        for t in range(12):
            merges.append(obj.tick())
            if t == 4:
                pos5 = obj.dev.get(F_POS)
        return merges, pos5
    merges_single, pos_c = run(single)
    ok_coll = True
    for mode in ("rows", "grid"):
        sh = ShardedAllPairsPhysics(dev, rank, world, collisions=mode)
        sh.load_system(conf, bd, bo)
        merges, pg = run(sh)
        same = merges == merges_single
        close = pg.shape == pos_c.shape and np.allclose(pg, pos_c, rtol=1e-11, atol=1e-9)
        if not (same and close) and rank == 0:
            print("  collision mode %s: merges equal %s, positions close %s" % (mode, same, close))
            print("   single:", merges_single)
            print("   sharded:", merges)
            if pg.shape == pos_c.shape:
                print("   max |dpos|", np.abs(pg - pos_c).max())
        ok_coll = ok_coll and same and close
    n_merges = sum(m is not None for m in merges_single)
    # push(): every rank uploads its block of rows, all-gathers complete the replicas
    sh.inner.pull("pos", "vel", "mass")
    sh.inner.pos += 3.0
    sh.inner.vel *= 0.5
    sh.push("mass", "pos", "vel")
    ok_coll = ok_coll and np.array_equal(sh.dev.get(F_POS), sh.inner.pos) and \
        np.array_equal(sh.dev.get(F_VEL), sh.inner.vel)
    # Kepler propagation: object blocks, no collective; each slice equals the single-GPU slice
    from volatilespace_b200._lib import KIND_VESSEL, K_EA, K_POS
    k = synth.kepler_objects(100003, seed=9)
    P = 32
    tt = np.linspace(-np.pi, np.pi, P)
    dev.kepler_load(KIND_VESSEL, P, k["a"], k["b"], k["f"], k["ecc"], k["pea"], k["n"], k["dr"],
                    k["ma"], k["ea"], k["ref"], tt)
    for _ in range(3):
        dev.kepler_tick(KIND_VESSEL, 2.0, k["body_pos"])
    ea_all = dev.kepler_get(KIND_VESSEL, K_EA, len(k["a"]))
    pos_all = dev.kepler_get(KIND_VESSEL, K_POS, len(k["a"]))
    sk = ShardedKepler(dev, KIND_VESSEL, P, k, rank, world, tt)
    for _ in range(3):
        sk.tick(2.0, k["body_pos"])
    ok_kep = np.array_equal(sk.get(K_EA), ea_all[sk.row0:sk.row1]) and \
        np.array_equal(sk.get(K_POS), pos_all[sk.row0:sk.row1])
    # MODE_REF tick: find_parents scan split over the ranks + MAX all-reduce; bit-identical
    from volatilespace_b200.dist import ShardedPhysics
    from volatilespace_b200.phys_editor import Physics
    from volatilespace_b200._lib import F_PARENTS
    e = synth.plummer2d(30000, seed=21)
    rng = np.random.default_rng(4)
    econf = {"gc": e["gc"], "rad_mult": 179.0, "coi_coef": 0.7}
    heavy = np.where(rng.uniform(0, 1, len(e["mass"])) < 0.01, 3000.0, 1.0)     # 1 % "planets" with moons
    ebd = {"mass": e["mass"] * heavy, "den": np.full(len(e["mass"]), 3000.0)}
    ebo = {"pos": e["pos"], "vel": e["vel"]}
    ref_single = Physics(dev)
    ref_single.load_system(econf, ebd, ebo)
    ref_sh = ShardedPhysics(Device(local), rank, world)
    ref_sh.load_system(econf, ebd, ebo)
    ok_ref_merges = True
    ref_single.body()
    ref_sh.body()
    n_ref_merges = 0
    for tick in range(8):
        # ticks 0-3: the sort-based candidate pass on every rank (default at this size);
        # ticks 4-7: the exhaustive scan, split over the ranks by row blocks.  Collision checks,
        # merges and mass edits run BETWEEN the gravity calls: they re-reserve the staging buffer
        # that holds best[] (the view the MAX all-reduce runs on must follow it)
        if tick == 4:
            os.environ["VSB_PARENTS"] = "scan"
        ref_single.gravity()
        ref_sh.gravity()
        ref_single.body()
        ref_sh.body()
        m_a, m_b = ref_single.inelastic_collision(), ref_sh.inelastic_collision()
        if m_a != m_b:
            ok_ref_merges = False
        n_ref_merges += m_a is not None
        if tick in (1, 5):
            for obj in (ref_single, ref_sh):
                obj.set_body_mass(17 + tick, float(obj.mass[17 + tick]) * 40.0)
    os.environ.pop("VSB_PARENTS", None)
    par = ref_single.dev.get(F_PARENTS)
    ok_ref = ok_ref_merges and np.array_equal(par, ref_sh.dev.get(F_PARENTS)) and \
        np.array_equal(ref_single.dev.get(F_POS), ref_sh.dev.get(F_POS)) and \
        np.array_equal(ref_single.dev.get(F_VEL), ref_sh.dev.get(F_VEL))
    n_children = int((par != 0).sum())
    flag = torch.tensor([int(ok), int(ok_host), int(ok_sym), int(ok_rep), int(ok_coll), int(ok_kep),
                         int(ok_ref)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multigpu_check n=%d world=%d ticks=%d: sharded==single %s, host path %s, "
              "sym within tolerance %s, sym replicas identical %s, collision ticks (%d merges) %s, "
              "kepler slices %s, MODE_REF ticks interleaved with %d merges and 2 mass edits (%d bodies whose "
              "parent is not body 0) %s" %
              (n, world, ticks, *[bool(x) for x in flag.tolist()[:4]], n_merges,
               bool(flag[4]), bool(flag[5]), n_ref_merges, n_children, bool(flag[6])))
    dist.barrier()
    dist.destroy_process_group()
    if not all(flag.tolist()):
        sys.exit(1)


if __name__ == "__main__":
    main()
