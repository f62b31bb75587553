"""CPU, world_size 2 over gloo: the host side of the row-sharded all-pairs tick
(volatilespace_b200.dist.partition / gather_rows) exchanges exactly the rows the
single-process step would produce.  The per-rank arithmetic is stood in for by the CPU
oracle (this is a test of the sharding and gather logic, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_rows_once():
    from volatilespace_b200.dist import partition
    for n in (1, 2, 15, 1023, 1024, 1025, 65536, 1048576, 262143):
        for world in (1, 2, 4, 8):
            rpr, ranges = partition(n, world)
            assert len(ranges) == world
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(r1 - r0 <= rpr for r0, r1 in ranges)
            npad = (n + 2047) // 2048 * 2048
            assert world * rpr <= npad          # the padded device array can hold the gather


def _worker(rank, world, port, n, steps, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from volatilespace_b200.dist import gather_rows, partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc.lib().vso_set_num_threads(2)
    rng = np.random.default_rng(3)
    mass = rng.uniform(0.5, 2.0, n)
    pos = rng.uniform(-1e3, 1e3, (n, 2))
    vel = rng.uniform(-1, 1, (n, 2))
    rpr, ranges = partition(n, world)
    r0, r1 = ranges[rank]
    npad = (n + 2047) // 2048 * 2048
    full = torch.zeros((npad, 2), dtype=torch.float64)
    full[:n] = torch.from_numpy(pos)
    v = vel[r0:r1].copy()
    for _ in range(steps):
        p = full[:n].numpy()
        acc = orc.allpairs_accel(mass, p, 1.0, r0, r1)
        v += acc
        full[r0:r1] = torch.from_numpy(p[r0:r1] + v)
        gather_rows(full, rpr, rank, world)
    if rank == 0:
        q.put(full[:n].numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [301, 512])
def test_sharded_tick_matches_single_process(n, oracle):
    import torch.multiprocessing as mp
    steps, world = 3, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    mass = rng.uniform(0.5, 2.0, n)
    pos = rng.uniform(-1e3, 1e3, (n, 2))
    vel = rng.uniform(-1, 1, (n, 2))
    want, _ = oracle.allpairs_step(mass, pos, vel, 1.0, steps)
    assert np.array_equal(got, want)


def _collision_worker(rank, world, port, n, q):
    """Row-split collision detection over gloo: every rank scans its row blocks (CPU oracle stands
    in for the kernel), one MIN all-reduce of the 64-bit key picks the pair."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from volatilespace_b200.dist import reduce_collision_key
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for case in range(3):
        rng = np.random.default_rng(40 + case)
        mass = rng.uniform(0.5, 2.0, n)
        pos = rng.uniform(-1e3, 1e3, (n, 2))
        rad = np.full(n, 2.0 if case < 2 else 1e-6)          # case 2: nothing collides
        # local minimum over the row blocks rank, rank+world, ... (blocks of 128 rows)
        key = -1
        for b0 in range(rank * 128, n - 1, world * 128):
            for i in range(b0, min(b0 + 128, n - 1)):
                d = np.sqrt((pos[i, 0] - pos[i + 1:, 0]) ** 2 + (pos[i, 1] - pos[i + 1:, 1]) ** 2)
                hit = np.flatnonzero(d <= rad[i] + rad[i + 1:])
                if len(hit):
                    k = (i << 32) | (i + 1 + int(hit[0]))
                    key = k if key < 0 else min(key, k)
                    break
            if key >= 0:
                break                                        # later blocks hold larger rows
        t = torch.tensor([key], dtype=torch.int64)
        reduce_collision_key(t)
        want = orc.check_collision(mass, pos, rad)
        k = int(t[0])
        got = (None, None) if k < 0 else (k >> 32, k & 0xffffffff)
        out.append((got, want))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_collision_key_allreduce(oracle):
    import torch.multiprocessing as mp
    world, n = 2, 700
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_collision_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (i, j), (body_del, body_add) in res:
        if i is None:
            assert body_del is None
        else:
            assert {i, j} == {body_del, body_add}       # the pair; del/add follows from the masses
    assert res[0][0][0] is not None and res[2][0][0] is None


def test_sharded_kepler_partition():
    """Object blocks of ShardedKepler cover every object once (host logic; no device needed)."""
    from volatilespace_b200.dist import partition
    for n, world in ((10, 4), (4194304, 8), (7, 8)):
        rpr, ranges = partition(n, world)
        covered = np.concatenate([np.arange(a, b) for a, b in ranges]) if n else np.zeros(0)
        assert np.array_equal(covered, np.arange(n))
