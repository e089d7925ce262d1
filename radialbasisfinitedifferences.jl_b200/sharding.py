"""Spatial-block sharding of the hot path across the GPUs of one box.

Every stencil is independent, so nodes are partitioned into contiguous slabs of the synthetic lattice (one slab
per rank, split along the slowest axis).  Each rank holds  [halo_lo | owned | halo_hi]  node coordinates: the
halo is generated locally from the closed-form node generator, so WEIGHT GENERATION NEEDS NO COMMUNICATION.
Column ids of the rank's operator rows are local ids into that layout.  Operator application needs one halo
exchange of the field per SpMV: the first/last `halo` owned values go to the neighbouring ranks
(torch.distributed P2P: NCCL send/recv over NVLink on the GPU box, gloo in the CPU tests).

The reference has no distributed code (src/domains/domains.jl:7-8 holds only commented-out includes); this is
the multi-GPU design of BASELINE.json's north_star.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class SlabShard:
    """Rank `rank` of `world` owns lattice rows [row0, row1) of a g^dim jittered lattice (linear ids x-fastest)."""
    rank: int
    world: int
    dim: int
    g: int
    halo_rows: int

    def __post_init__(self):
        # A halo of `halo_rows` lattice rows is filled by ONE hop from the adjacent rank, so every rank must own at least
        # that many rows; thinner slabs would make a neighbour ship rows out of its own (stale) halo.
        if self.world > 1 and self.halo_rows > 0:
            thinnest = min((self.g * (r + 1)) // self.world - (self.g * r) // self.world for r in range(self.world))
            if thinnest < self.halo_rows:
                raise ValueError(f"SlabShard: the thinnest slab owns {thinnest} lattice rows < halo_rows = {self.halo_rows}; "
                                 "use fewer ranks, a larger lattice or a narrower halo (single-hop halo exchange)")

    @property
    def row_size(self) -> int:          # nodes per slowest-axis row (a line in 2-D, a plane in 3-D)
        return self.g ** (self.dim - 1)

    @property
    def row0(self) -> int:
        return (self.g * self.rank) // self.world

    @property
    def row1(self) -> int:
        return (self.g * (self.rank + 1)) // self.world

    @property
    def lo_rows(self) -> int:           # halo rows actually present below / above (domain boundary: none)
        return min(self.halo_rows, self.row0)

    @property
    def hi_rows(self) -> int:
        return min(self.halo_rows, self.g - self.row1)

    @property
    def n_owned(self) -> int:
        return (self.row1 - self.row0) * self.row_size

    @property
    def n_lo(self) -> int:
        return self.lo_rows * self.row_size

    @property
    def n_hi(self) -> int:
        return self.hi_rows * self.row_size

    @property
    def n_local(self) -> int:
        return self.n_lo + self.n_owned + self.n_hi

    @property
    def first_local_id(self) -> int:    # global linear id of local node 0
        return (self.row0 - self.lo_rows) * self.row_size

    @property
    def first_owned_id(self) -> int:
        return self.row0 * self.row_size

    def missing_edges(self):
        """Slowest-axis coordinates beyond which nodes are NOT present locally: (below, above); None at a domain edge.
        A lattice row j holds coordinates in [(j+0.25)/g, (j+0.75)/g]."""
        below = None if self.row0 - self.lo_rows == 0 else (self.row0 - self.lo_rows - 1 + 0.75) / self.g
        above = None if self.row1 + self.hi_rows == self.g else (self.row1 + self.hi_rows + 0.25) / self.g
        return below, above

    def halo_is_sufficient(self, x_last_owned, kth_dist2):
        """Exactness check: every owned stencil's farthest neighbour is strictly closer than any node that is not
        held locally.  x_last_owned: slowest-axis coordinate of the owned nodes, kth_dist2: their largest squared
        neighbour distance (torch tensors or NumPy arrays)."""
        below, above = self.missing_edges()
        ok = True
        if below is not None:
            gap = x_last_owned - below
            ok = ok and bool(((gap > 0) & (gap * gap > kth_dist2)).all())
        if above is not None:
            gap = above - x_last_owned
            ok = ok and bool(((gap > 0) & (gap * gap > kth_dist2)).all())
        return ok


class _HaloWork:
    """handle of an in-flight halo exchange: wait() makes the current stream (or the host under gloo) wait for it"""

    def __init__(self, reqs):
        self.reqs = reqs

    def wait(self):
        for r in self.reqs:
            r.wait()


def exchange_halo(u_local, shard: SlabShard, group=None, async_op=False):
    """Fill the halo parts of u_local = [halo_lo | owned | halo_hi] (1-D torch tensor) from the neighbouring ranks.
    The owned part must be current.  One batched send/recv group per call (ncclGroupStart/End under NCCL).
    async_op=True returns a handle immediately so that rows which do not touch the halo can be applied while the
    exchange is in flight (NCCL runs it on its own stream); call .wait() before the boundary rows."""
    import torch.distributed as dist
    s = shard
    if s.world == 1:
        return _HaloWork([]) if async_op else u_local
    ops = []
    o0, o1 = s.n_lo, s.n_lo + s.n_owned
    if s.rank > 0:
        # rank-1 keeps min(halo_rows, rows it lacks above) rows of ours as its halo_hi: exactly our first rows
        below = SlabShard(s.rank - 1, s.world, s.dim, s.g, s.halo_rows)
        ops.append(dist.P2POp(dist.isend, u_local[o0:o0 + below.n_hi], s.rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, u_local[0:s.n_lo], s.rank - 1, group))
    if s.rank < s.world - 1:
        above = SlabShard(s.rank + 1, s.world, s.dim, s.g, s.halo_rows)
        ops.append(dist.P2POp(dist.isend, u_local[o1 - above.n_lo:o1], s.rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, u_local[o1:o1 + s.n_hi], s.rank + 1, group))
    work = _HaloWork(dist.batch_isend_irecv(ops))
    if async_op:
        return work
    work.wait()
    return u_local


def boundary_row_ranges(shard: SlabShard):
    """Owned rows split into (low boundary, interior, high boundary) half-open ranges.  The split is geometric (rows within
    halo_rows lattice rows of a slab face); that the interior rows really reference no halo column is NOT implied by
    halo_is_sufficient -- check it with interior_rows_are_halo_free(colind, shard) before overlapping the interior rows
    with a halo exchange."""
    s = shard
    lo = min(s.n_owned, s.n_lo)                      # n_lo == lo_rows * row_size (0 at the domain edge)
    hi = min(s.n_owned - lo, s.n_hi)
    return (0, lo), (lo, s.n_owned - hi), (s.n_owned - hi, s.n_owned)


def interior_rows_are_halo_free(colind, shard: SlabShard) -> bool:
    """True when no row of the interior range of boundary_row_ranges references a halo column, i.e. when those rows may be
    applied while neighbours are still storing into the halo.  colind: [n_owned, n] local column ids (torch or NumPy)."""
    (i0, i1) = boundary_row_ranges(shard)[1]
    if i1 <= i0:
        return True
    ci = colind[i0:i1]
    return bool(int(ci.min()) >= shard.n_lo and int(ci.max()) < shard.n_lo + shard.n_owned)


class PeerHalo:
    """NVLink peer-memory halo exchange (csrc/halo.cu): the field lives in a CUDA-IPC buffer, the neighbours store their
    boundary values straight into its halo regions and publish an epoch flag; no NCCL call on the data path.

        halo = PeerHalo(ctx, shard)          # collective: exchanges IPC handles through torch.distributed
        u = halo.field                       # torch view [halo_lo | owned | halo_hi] of the IPC buffer
        halo.push(); <interior rows>; halo.wait(); <boundary rows>; halo.ack()
    """

    def __init__(self, ctx, shard: SlabShard, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._lib import Halo
        self.ctx, self.shard, self.epoch = ctx, shard, 0
        s = shard
        nbytes = s.n_local * 8 + 64
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        ctx._check(ctx._L.rbffd_ipc_alloc(ctx._h, nbytes, C.byref(ptr), handle))
        self._ptr = ptr.value
        self._peers = []
        handles = [None] * s.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)

        def open_peer(r):
            p = C.c_void_p()
            ctx._check(ctx._L.rbffd_ipc_open(ctx._h, handles[r], C.byref(p)))
            self._peers.append(p.value)
            return p.value

        h = Halo()
        h.u, h.n_lo, h.n_owned, h.n_hi = self._ptr, s.n_lo, s.n_owned, s.n_hi
        h.flags = self._ptr + s.n_local * 8
        if s.rank > 0:
            below = SlabShard(s.rank - 1, s.world, s.dim, s.g, s.halo_rows)
            base = open_peer(s.rank - 1)
            h.peer_lo_u, h.peer_lo_flags = base, base + below.n_local * 8
            h.peer_lo_offset, h.count_to_lo = below.n_lo + below.n_owned, below.n_hi
        if s.rank < s.world - 1:
            above = SlabShard(s.rank + 1, s.world, s.dim, s.g, s.halo_rows)
            base = open_peer(s.rank + 1)
            h.peer_hi_u, h.peer_hi_flags = base, base + above.n_local * 8
            h.peer_hi_offset, h.count_to_hi = 0, above.n_lo
        self._h = h

        class _Raw:          # zero-copy torch view of the IPC buffer
            __cuda_array_interface__ = {"shape": (s.n_local,), "typestr": "<f8", "data": (self._ptr, False), "version": 3, "strides": None}
        self._raw = _Raw()
        self.field = torch.as_tensor(self._raw, device=torch.device("cuda", ctx.device))
        dist.barrier(group=group)

    def push(self):
        import ctypes as C
        self.epoch += 1
        self.ctx._check(self.ctx._L.rbffd_halo_push_device(self.ctx._h, C.byref(self._h), self.epoch))

    def wait(self):
        import ctypes as C
        self.ctx._check(self.ctx._L.rbffd_halo_wait_device(self.ctx._h, C.byref(self._h), self.epoch))

    def ack(self):
        import ctypes as C
        self.ctx._check(self.ctx._L.rbffd_halo_ack_device(self.ctx._h, C.byref(self._h), self.epoch))

    def close(self, group=None):
        # collective: a neighbour's push / ack kernel may still be storing into this buffer
        import torch.distributed as dist
        self.ctx.synchronize()
        if dist.is_initialized():
            dist.barrier(group=group)
        for p in self._peers:
            self.ctx._L.rbffd_ipc_close(self.ctx._h, p)
        self._peers = []
        if self._ptr:
            self.field = None
            self.ctx._L.rbffd_ipc_free(self.ctx._h, self._ptr)
            self._ptr = None
