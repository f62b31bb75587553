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
