"""Row sharding of the all-pairs gravity step across the GPUs of one node
(SURVEY.md 8e): one process per GPU, target rows split in contiguous blocks,
every rank keeps full mass/position replicas, one all-gather of the updated
positions per tick over NCCL (NVLink/NVSwitch).  The gather is 16 B/body/tick
(16 MiB at 1M bodies) against ~0.1-1 s of FP64 work per tick, so a plain NCCL
all-gather on the kernel's own stream is all the exchange this path needs.

torch is used only for torch.distributed and to view the library's device
buffers as tensors (zero copy); no torch kernel touches the data.
"""
import numpy as np

from ._lib import F_MASS, F_POS, F_VEL


def partition(n, world):
    """Contiguous row blocks of ceil(n/world) rows (the last ranks may be short or
    empty).  Returns (rows_per_rank, [(row0,row1)] * world)."""
    rpr = (n + world - 1) // world
    return rpr, [(min(r * rpr, n), min((r + 1) * rpr, n)) for r in range(world)]


def gather_rows(full, rpr, rank, world, group=None):
    """In-place all-gather of row blocks: rank r contributes full[r*rpr:(r+1)*rpr].
    `full` is a (>= world*rpr, 2) torch tensor (CPU for gloo, CUDA for NCCL)."""
    import torch.distributed as dist
    if world == 1:
        return
    out = full[: world * rpr]
    mine = full[rank * rpr:(rank + 1) * rpr]
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_gather_into_tensor(out, mine, group=group)      # in place (NCCL allows it)
    else:
        parts = [out[r * rpr:(r + 1) * rpr].clone() for r in range(world)]
        dist.all_gather(parts, mine.clone(), group=group)
        for r in range(world):
            out[r * rpr:(r + 1) * rpr] = parts[r]


def reduce_collision_key(key, group=None):
    """MIN over the ranks of a collision key (i << 32 | j held as int64; -1 = all ones = "no
    collision"), in place on a 1-element int64 tensor (CPU for gloo, CUDA for NCCL).  The
    sentinel is mapped to INT64_MAX around the reduction so that it loses against any pair."""
    import torch
    import torch.distributed as dist
    big = torch.iinfo(torch.int64).max
    tmp = torch.where(key < 0, torch.full_like(key, big), key)
    dist.all_reduce(tmp, op=dist.ReduceOp.MIN, group=group)
    key.copy_(torch.where(tmp == big, torch.full_like(tmp, -1), tmp))
    return key


class _CudaView:
    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def device_tensor(dev, field, rows, width=2):
    """Zero-copy torch view of a library-owned device array."""
    import torch
    shape = (rows, width) if width > 1 else (rows,)
    return torch.as_tensor(_CudaView(dev.device_ptr(field), shape), device="cuda:%d" % dev.index)


class ShardedAllPairs:
    """MODE_ALLPAIRS over `world` ranks.  Results are bit-identical to the single-GPU
    step: every rank forms its rows' sums in the same fixed order."""

    def __init__(self, dev, mass, pos, vel, gc, rank=0, world=1, group=None):
        self.dev, self.gc, self.rank, self.world, self.group = dev, float(gc), rank, world, group
        self.n = len(mass)
        self.rpr, self.ranges = partition(self.n, world)
        self.row0, self.row1 = self.ranges[rank]
        dev.upload_bodies(mass, pos, vel)
        dev.set_gravity_mode("rows")      # the shard-invariant kernel (auto would pick sym at world 1)
        dev.set_rows(self.row0, self.row1)
        self._stream = None
        self._pos_t = None
        if world > 1:
            import torch
            npad = (max(self.n, 1) + 2047) // 2048 * 2048     # VSB_PAD of the library
            assert world * self.rpr <= npad
            self._pos_t = device_tensor(dev, F_POS, npad)
            self._mass_t = device_tensor(dev, F_MASS, npad, width=1)
            self._stream = torch.cuda.ExternalStream(dev.stream_handle(),
                                                     device="cuda:%d" % dev.index)

    def tick(self):
        """accel + integrate own rows, then all-gather positions; asynchronous."""
        self.dev.allpairs_tick_async(self.gc)
        if self.world > 1:
            import torch
            with torch.cuda.stream(self._stream):      # NCCL orders itself after our kernels
                gather_rows(self._pos_t, self.rpr, self.rank, self.world, self.group)

    def tick_host(self, mass, pos, vel):
        """End-to-end tick over HOST buffers: each rank uploads ITS rows of mass / pos / vel,
        the replicas of mass and pos are completed by NCCL all-gathers over NVLink, and the rank
        downloads its own rows of pos / vel (updated in place)."""
        d, r0, r1 = self.dev, self.row0, self.row1
        if r1 > r0:
            d.set(F_MASS, mass[r0:r1], first=r0)
            d.set(F_POS, pos[r0:r1], first=r0)
            d.set(F_VEL, vel[r0:r1], first=r0)
        if self.world > 1:
            import torch
            with torch.cuda.stream(self._stream):
                gather_rows(self._mass_t, self.rpr, self.rank, self.world, self.group)
                gather_rows(self._pos_t, self.rpr, self.rank, self.world, self.group)
        self.tick()
        if self.world > 1:
            self._stream.synchronize()
        if r1 > r0:
            d.get(F_POS, first=r0, count=r1 - r0, out=pos[r0:r1])
            d.get(F_VEL, first=r0, count=r1 - r0, out=vel[r0:r1])

    def host_bytes_per_tick(self):
        rows = self.row1 - self.row0
        return rows * (8 + 16 + 16), rows * 32

    def positions(self):
        self.dev.sync()
        return self.dev.get(F_POS)

    def velocities_own(self):
        self.dev.sync()
        return self.dev.get(F_VEL, first=self.row0, count=self.row1 - self.row0)


class ShardedAllPairsSym:
    """MODE_ALLPAIRS with the symmetric kernel (every unordered pair once) over `world`
    ranks: each rank evaluates its share of the unordered block pairs, the partial
    accelerations are summed with one NCCL all-reduce (16 B/body), and every rank integrates
    all bodies (the all-reduce result is identical on all ranks, so the replicas stay in
    lock-step).  Deterministic run to run for a given world size."""

    def __init__(self, dev, mass, pos, vel, gc, rank=0, world=1, group=None):
        self.dev, self.gc, self.rank, self.world, self.group = dev, float(gc), rank, world, group
        self.n = len(mass)
        self.rpr, self.ranges = partition(self.n, world)
        self.row0, self.row1 = self.ranges[rank]       # the slice of HOST data this rank owns
        dev.upload_bodies(mass, pos, vel)
        dev.set_gravity_mode("sym")
        self._stream = None
        if world > 1:
            import torch
            from ._lib import F_ACC
            npad = (max(self.n, 1) + 2047) // 2048 * 2048     # VSB_PAD of the library
            assert world * self.rpr <= npad
            self._acc_t = device_tensor(dev, F_ACC, self.n)
            self._pos_t = device_tensor(dev, F_POS, npad)
            self._vel_t = device_tensor(dev, F_VEL, npad)
            self._mass_t = device_tensor(dev, F_MASS, npad, width=1)
            self._stream = torch.cuda.ExternalStream(dev.stream_handle(),
                                                     device="cuda:%d" % dev.index)

    def tick(self):
        if self.world == 1:
            self.dev.allpairs_tick_async(self.gc)
            return
        import torch
        import torch.distributed as dist
        self.dev.allpairs_sym_partial_async(self.gc, self.rank, self.world)
        with torch.cuda.stream(self._stream):
            dist.all_reduce(self._acc_t, group=self.group)
        self.dev.allpairs_integrate_async()

    def tick_host(self, mass, pos, vel):
        """End-to-end tick over host buffers.  Each rank moves only ITS slice of the host
        arrays over PCIe (rows [row0,row1) of mass / pos / vel up, of pos / vel down); the
        device replicas are completed by NCCL all-gathers over NVLink."""
        d, r0, r1 = self.dev, self.row0, self.row1
        if r1 > r0:
            d.set(F_MASS, mass[r0:r1], first=r0)
            d.set(F_POS, pos[r0:r1], first=r0)
            d.set(F_VEL, vel[r0:r1], first=r0)
        if self.world > 1:
            import torch
            with torch.cuda.stream(self._stream):
                for t in (self._mass_t, self._pos_t, self._vel_t):
                    gather_rows(t, self.rpr, self.rank, self.world, self.group)
        self.tick()
        if self.world > 1:
            self._stream.synchronize()
        if r1 > r0:
            d.get(F_POS, first=r0, count=r1 - r0, out=pos[r0:r1])
            d.get(F_VEL, first=r0, count=r1 - r0, out=vel[r0:r1])

    def host_bytes_per_tick(self):
        rows = self.row1 - self.row0
        return rows * (8 + 16 + 16), rows * 32

    def positions(self):
        self.dev.sync()
        return self.dev.get(F_POS)


class ShardedAllPairsPhysics:
    """The MODE_ALLPAIRS editor tick -- gravity(); body(); inelastic_collision()
    (volatilespace_b200.phys_editor.AllPairsPhysics) -- over `world` ranks:
      * gravity: symmetric kernel, block-pair tasks dealt to the ranks, one all-reduce of the
        partial accelerations, every rank integrates all bodies (replicas stay identical);
      * collision detect: `collisions="rows"` splits the exhaustive scan by row blocks, each rank
        finds its local minimum key (i << 32 | j) and one 64-bit MIN all-reduce picks the pair;
        `collisions="grid"` (default) runs the sort-based candidate pass on every rank -- it takes
        ~0.1 ms at 256k bodies, less than the latency of a collective, and the replicas are
        identical so the answers are too;
      * merge: applied redundantly (deterministically) on every rank; the task table follows the
        shrinking body count.
    Every rank holds the full host mirrors, like the single-GPU class."""

    def __init__(self, device, rank=0, world=1, group=None, collisions="grid"):
        from .phys_editor import AllPairsPhysics
        assert collisions in ("grid", "rows")
        self.inner = AllPairsPhysics(device)
        self.dev, self.rank, self.world, self.group = self.inner.dev, rank, world, group
        self.collisions = collisions
        self._views_n = -1
        self._stream = None
        self.dev.set_gravity_mode("sym")

    def load_system(self, conf, body_data, body_orb_data):
        self.inner.load_system(conf, body_data, body_orb_data)
        self._views_n = -1

    def __getattr__(self, name):           # mass, pos, vel, rad, den, gc, ... of the inner class
        return getattr(self.inner, name)

    def _views(self):
        """torch views of the acceleration array and the collision key word; rebuilt whenever the
        library's addresses or the body count changed (a resize may move the body slab)."""
        import torch
        from ._lib import F_ACC
        n = len(self.inner.mass)
        key = (self.dev.device_ptr(F_ACC), self.dev.collision_key_ptr(), n)
        if self._views_n != key:
            self._acc_t = device_tensor(self.dev, F_ACC, n)
            kview = _CudaView(key[1], (1,), "<i8")
            self._key_t = torch.as_tensor(kview, device="cuda:%d" % self.dev.index)
            self._views_n = key
        if self._stream is None:
            self._stream = torch.cuda.ExternalStream(self.dev.stream_handle(),
                                                     device="cuda:%d" % self.dev.index)
        return self._acc_t, self._key_t

    def gravity(self):
        if self.world == 1:
            return self.inner.gravity()
        import torch
        import torch.distributed as dist
        acc, _ = self._views()
        self.dev.allpairs_sym_partial_async(self.inner.gc, self.rank, self.world)
        with torch.cuda.stream(self._stream):
            dist.all_reduce(acc, group=self.group)
        self.dev.allpairs_integrate_async()
        self.inner._touched("pos", "vel")

    def push(self, *names):
        """Host mirrors -> device for the named fields ("mass", "pos", "vel", ...): every rank holds
        the full mirrors, so each uploads only ITS block of rows over PCIe and the device replicas
        are completed by in-place NCCL all-gathers over NVLink (same result as inner._push on every
        rank, 1/world of the PCIe traffic)."""
        inner = self.inner
        if self.world == 1:
            return inner._push(*names)
        import torch
        from .phys_editor import _FIELDS, _WIDE
        n = len(inner.mass)
        rpr, ranges = partition(n, self.world)
        r0, r1 = ranges[self.rank]
        npad = (max(n, 1) + 2047) // 2048 * 2048
        self._views()
        for name in names:
            assert name != "parents", "push() moves the float64 fields"
            arr = getattr(inner, name)
            if r1 > r0:
                self.dev.set(_FIELDS[name], arr[r0:r1], first=r0)
            view = device_tensor(self.dev, _FIELDS[name], npad, width=2 if name in _WIDE else 1)
            with torch.cuda.stream(self._stream):
                gather_rows(view, rpr, self.rank, self.world, self.group)
            inner._stale.discard(name)
        self._stream.synchronize()

    def body(self):
        self.inner.body()

    def check_collision(self):
        if self.world == 1 or self.collisions == "grid":
            return self.dev.check_collision()
        import torch
        _, key = self._views()
        self.dev.check_collision_part_async(self.rank, self.world)
        with torch.cuda.stream(self._stream):
            reduce_collision_key(key, self.group)
        self._stream.synchronize()
        return self.dev.check_collision_finish()

    def inelastic_collision(self):
        inner = self.inner
        inner.check_collision = self.check_collision      # same merge code, sharded detection
        try:
            return inner.inelastic_collision()
        finally:
            del inner.check_collision

    def tick(self):
        self.gravity()
        self.body()
        return self.inelastic_collision()


class ShardedKepler:
    """Game-mode propagation + orbit curves over `world` ranks: contiguous object
    blocks, the (small) set of parent bodies replicated on every rank, no collective -- each
    object's tick depends only on its own elements and its parent's position."""

    def __init__(self, dev, kind, P, elements, rank=0, world=1, t=None):
        """elements: dict of full-length arrays a b f ecc pea n dr ma ea ref (the reference's
        orbit arrays after calc_orb_one); this rank keeps objects [row0, row1)."""
        self.dev, self.kind, self.P, self.rank, self.world = dev, kind, int(P), rank, world
        self.n = len(elements["a"])
        self.rpr, self.ranges = partition(self.n, world)
        self.row0, self.row1 = self.ranges[rank]
        sl = slice(self.row0, self.row1)
        self.count = self.row1 - self.row0
        if self.count:
            dev.kepler_load(kind, self.P, *(np.ascontiguousarray(elements[k][sl]) for k in (
                "a", "b", "f", "ecc", "pea", "n", "dr", "ma", "ea", "ref")), t)

    def tick(self, warp, body_pos):
        """move + curves of this rank's objects (results stay on its device)."""
        if self.count:
            self.dev.kepler_tick(self.kind, warp, body_pos)

    def tick_async(self, warp):
        if self.count:
            self.dev.kepler_tick_async(self.kind, warp)

    def get(self, field):
        """This rank's slice of a state field (objects [row0, row1))."""
        return self.dev.kepler_get(self.kind, field, self.count)


class ShardedPhysics:
    """The MODE_REF editor tick (phys_editor.Physics.gravity, phys_editor.py:357-378) over `world`
    ranks: the O(N^2) find_parents scan is split by row blocks of the mass-sorted
    bodies, one MAX all-reduce of the int32 result array (4 B/body) completes it, and every rank
    runs the O(N) remainder (gravity_one per body, chain composition, integrator) on its full
    replica.  Integer results: bit-identical to the single-GPU tick."""

    def __init__(self, device, rank=0, world=1, group=None):
        from .phys_editor import Physics
        self.inner = Physics(device)
        self.dev, self.rank, self.world, self.group = self.inner.dev, rank, world, group
        self._best_key = None
        self._stream = None

    def __getattr__(self, name):
        return getattr(self.inner, name)

    def _best(self):
        """torch view of the library's best[] array.  best[] lives in a staging buffer that other
        passes (the collision candidate pass, simplified_orbit_coi) may re-allocate between ticks,
        so the address is asked for on EVERY tick -- after gravity_ref_begin, which reserves the
        buffer -- and the view is rebuilt whenever the address or the body count changed."""
        import torch
        n = len(self.inner.mass)
        key = (self.dev.find_parents_best_ptr(), n)
        if self._best_key != key:
            view = _CudaView(key[0], (n,), "<i4")
            self._best_t = torch.as_tensor(view, device="cuda:%d" % self.dev.index)
            self._best_key = key
        if self._stream is None:
            self._stream = torch.cuda.ExternalStream(self.dev.stream_handle(),
                                                     device="cuda:%d" % self.dev.index)
        return self._best_t

    def gravity(self):
        inner = self.inner
        if self.world == 1 or len(inner.mass) <= 1024:     # small systems: the fused one-CTA tick
            return inner.gravity()
        import torch
        import torch.distributed as dist
        order = None if inner._order_on_device else inner._order()
        self.dev.gravity_ref_begin_async(order, self.rank, self.world)
        best = self._best()          # after begin: begin may have re-allocated the staging buffer
        with torch.cuda.stream(self._stream):
            dist.all_reduce(best, op=dist.ReduceOp.MAX, group=self.group)
        self.dev.gravity_ref_finish(inner.gc)
        inner._order_on_device = True
        inner._touched("parents", "rel_vel", "vel", "pos")
