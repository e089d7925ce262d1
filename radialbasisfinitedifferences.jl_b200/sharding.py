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


# ---------------------------------------------------------------------------------------------------------------------
# General spatial-block shards (csrc/shard.cu): arbitrary node sets, k-way coordinate-quantile blocks (2x2x2 on 8 GPUs),
# halo = stencil closure, local numbering [interior | boundary | halo], halo exchange FUSED into the SpMV launch.
# ---------------------------------------------------------------------------------------------------------------------
def plan(X, nparts: int, blocks=None):
    """Partition the nodes X [N, d] into `nparts` spatial blocks (rbffd_shard_plan_host; host only, no device needed).
    blocks = (b0, b1[, b2]) fixes the block grid; None picks the factorisation with the smallest cut surface.
    Returns part[N] int32 in [0, nparts)."""
    import ctypes as C
    import numpy as np
    from . import _lib
    X = np.ascontiguousarray(X, np.float64)
    N, dim = X.shape
    part = np.empty(N, np.int32)
    b = None
    if blocks is not None:
        b = (C.c_int32 * 3)(*(list(blocks) + [1] * (3 - len(blocks))))
    rc = _lib.lib().rbffd_shard_plan_host(X.ctypes.data, N, dim, nparts, b, part.ctypes.data)
    if rc != 0:
        raise _lib.RbffdError(rc, f"rbffd_shard_plan_host(N={N}, dim={dim}, nparts={nparts}, blocks={blocks})")
    return part


class Shard:
    """One rank's block of a sharded node set: owned nodes + halo (stencil closure), the stencils of the owned nodes in local
    numbering, and the wiring of the halo exchange.  Generation needs no communication; rows come back in LOCAL order
    [interior | boundary], `global_ids()` maps local ids (owned, then halo) to the caller's numbering."""

    def __init__(self, ctx, handle, rank, nparts, n):
        import ctypes as C
        self.ctx, self._h, self.rank, self.nparts, self.n = ctx, handle, rank, nparts, n
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        ctx._check(ctx._L.rbffd_shard_info(handle, C.byref(a), C.byref(b), C.byref(c)))
        self.n_owned, self.n_interior, self.n_halo = a.value, b.value, c.value
        x, s = C.c_void_p(), C.c_void_p()
        ctx._check(ctx._L.rbffd_shard_device_arrays(handle, C.byref(x), C.byref(s)))
        self.X_local_ptr, self.stencils_ptr = x.value, s.value
        self._bufs = []

    @classmethod
    def from_host(cls, ctx, X, part, nparts, rank, n):
        """shard of `rank` from the full node set X [N, d] and its partition (host arrays)"""
        import ctypes as C
        import numpy as np
        X = np.ascontiguousarray(X, np.float64)
        part = np.ascontiguousarray(part, np.int32)
        h = C.c_void_p()
        ctx._check(ctx._L.rbffd_shard_create_host(ctx._h, X.ctypes.data, X.shape[0], X.shape[1], part.ctypes.data, nparts, rank, n, C.byref(h)))
        return cls(ctx, h, rank, nparts, n)

    @classmethod
    def from_device(cls, ctx, dim, n, Xc_ptr, gid_ptr, owner_ptr, nc, box_lo, box_hi, rank, nparts):
        """shard from device-resident candidates in ASCENDING global id; box = region all of whose nodes are candidates"""
        import ctypes as C
        lo = (C.c_double * 3)(*(list(box_lo) + [0.0] * (3 - len(box_lo))))
        hi = (C.c_double * 3)(*(list(box_hi) + [0.0] * (3 - len(box_hi))))
        h = C.c_void_p()
        ctx._check(ctx._L.rbffd_shard_create_device(ctx._h, dim, n, Xc_ptr, gid_ptr, owner_ptr, nc, lo, hi, rank, nparts, C.byref(h)))
        return cls(ctx, h, rank, nparts, n)

    @classmethod
    def lattice_block(cls, ctx, dim, g, seed, blocks, rank, n, margin=None):
        """Block `rank` of the closed-form jittered lattice under a regular blocks[dim] block grid: the
        candidates (own block + `margin` lattice cells around it) are generated on the device, nothing is communicated.
        A node of cell c has coordinates in [(c + 0.25)/g, (c + 0.75)/g], which gives the box of the exactness proof."""
        import ctypes as C
        import math
        import torch
        from . import _lib
        blocks = list(blocks)
        nparts = math.prod(blocks)
        b, r = [0] * dim, rank
        for a in reversed(range(dim)):
            b[a] = r % blocks[a]
            r //= blocks[a]
        if margin is None:
            vd = math.pi if dim == 2 else 4.0 * math.pi / 3.0
            margin = int(math.ceil(1.35 * (n / vd) ** (1.0 / dim))) + 1
        dev = torch.device("cuda", ctx.device)
        while True:
            lo, hi, blo, bhi = [], [], [], []
            for a in range(dim):
                c0 = -((-b[a] * g) // blocks[a])                 # first cell with (c * B) // g == b
                c1 = -((-(b[a] + 1) * g) // blocks[a])
                l, h = max(0, c0 - margin), min(g, c1 + margin)
                lo.append(l); hi.append(h)
                blo.append((l - 0.25) / g if l > 0 else -math.inf)
                bhi.append((h + 0.25) / g if h < g else math.inf)
            nc = math.prod(h - l for l, h in zip(lo, hi))
            Xc = torch.empty(nc * dim, dtype=torch.float64, device=dev)
            gid = torch.empty(nc, dtype=torch.int64, device=dev)
            own = torch.empty(nc, dtype=torch.int32, device=dev)
            L = (C.c_int64 * 3)(*(lo + [0] * (3 - dim)))
            H = (C.c_int64 * 3)(*(hi + [0] * (3 - dim)))
            B = (C.c_int32 * 3)(*(blocks + [1] * (3 - dim)))
            ctx._check(ctx._L.rbffd_jittered_lattice_box_device(ctx._h, dim, g, seed, L, H, B, Xc.data_ptr(), gid.data_ptr(), own.data_ptr()))
            try:
                s = cls.from_device(ctx, dim, n, Xc.data_ptr(), gid.data_ptr(), own.data_ptr(), nc, blo, bhi, rank, nparts)
            except _lib.RbffdError as e:
                if e.code != _lib.ERR_HALO or all(l == 0 and h == g for l, h in zip(lo, hi)):
                    raise
                margin += 1
                continue
            ctx.synchronize()
            s.margin = margin
            return s

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._L.rbffd_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self):
        """synchronise and raise RbffdError(ERR_HALO) if a fused halo exchange timed out waiting for a peer"""
        self.ctx._check(self.ctx._L.rbffd_shard_status(self._h))

    def global_ids(self, index_base=0):
        import numpy as np
        out = np.empty(self.n_owned + self.n_halo, np.int64)
        self.ctx._check(self.ctx._L.rbffd_shard_global_ids_host(self._h, index_base, out.ctypes.data))
        return out

    def generate(self, p, polydeg, ops, dim, kernel=0):
        """Operators of the owned rows over the local pattern (rows n_owned, columns n_owned + n_halo): one fused weight
        solve per owned node, no communication.  Returns an Operator whose value planes follow `ops`."""
        import torch
        from .api import make_options
        opts = make_options(dim, p, self.n, polydeg, ops, kernel=kernel)
        dev = torch.device("cuda", self.ctx.device)
        colind = torch.empty(self.n_owned * self.n, dtype=torch.int32, device=dev)
        vals = torch.empty(len(ops) * self.n_owned * self.n, dtype=torch.float64, device=dev)
        self.ctx.weights_device(opts, self.X_local_ptr, self.n_owned + self.n_halo, self.stencils_ptr, colind.data_ptr(), vals.data_ptr(),
                                Y_ptr=self.X_local_ptr, M=self.n_owned, NS=self.n_owned)
        op = self.ctx.operator_from_device(self.n_owned, self.n_owned + self.n_halo, self.n, len(ops), colind.data_ptr(), vals.data_ptr())
        op._keep = (colind, vals)          # the operator borrows these buffers
        return op

    # ---- wiring ----
    def recv_count(self, peer):
        import ctypes as C
        c = C.c_int64()
        self.ctx._check(self.ctx._L.rbffd_shard_recv_count(self._h, peer, C.byref(c)))
        return c.value

    def send_count(self, peer):
        import ctypes as C
        c = C.c_int64()
        self.ctx._check(self.ctx._L.rbffd_shard_send_count(self._h, peer, C.byref(c)))
        return c.value

    def recv_ids(self, peer):
        import numpy as np
        out = np.empty(self.recv_count(peer), np.int64)
        self.ctx._check(self.ctx._L.rbffd_shard_recv_ids_host(self._h, peer, 0, out.ctypes.data))
        return out

    def set_send_ids(self, peer, ids):
        import numpy as np
        ids = np.ascontiguousarray(ids, np.int64)
        self.ctx._check(self.ctx._L.rbffd_shard_set_send_ids_host(self._h, peer, 0, ids.ctypes.data, ids.size))

    def finalize(self):
        import ctypes as C
        import numpy as np
        handle = C.create_string_buffer(64)
        offsets = np.empty(2 * self.nparts + 1, np.int64)
        self.ctx._check(self.ctx._L.rbffd_shard_finalize(self._h, handle, offsets.ctypes.data))
        return bytes(handle.raw), offsets

    def connect(self, peer, handle, fwd_offset, rev_offset, flags_offset):
        self.ctx._check(self.ctx._L.rbffd_shard_connect(self._h, peer, handle, int(fwd_offset), int(rev_offset), int(flags_offset)))

    def wire(self, group=None, ipc=True):
        """Collective over the ranks of `group` (rank r of the group = shard r): ships every rank's halo requests to the
        owners, allocates the CUDA-IPC inboxes and maps the peers' inboxes (NVLink peer stores).  torch.distributed is the
        plumbing; a Julia host would do the same three steps over MPI.  ipc=False stops after the inbox allocation: the
        halo values then travel through exchange_collective (no peer memory is mapped)."""
        import numpy as np
        import torch.distributed as dist
        want = {p: self.recv_ids(p) for p in range(self.nparts) if p != self.rank and self.recv_count(p) > 0}
        allwant = [None] * self.nparts
        dist.all_gather_object(allwant, want, group=group)
        for p in range(self.nparts):
            if p != self.rank and self.rank in allwant[p]:
                self.set_send_ids(p, allwant[p][self.rank])
        mine = self.finalize()
        if not ipc:
            dist.barrier(group=group)
            return
        everyone = [None] * self.nparts
        dist.all_gather_object(everyone, mine, group=group)
        for p in range(self.nparts):
            if p == self.rank or (self.recv_count(p) == 0 and self.send_count(p) == 0):
                continue
            handle, off = everyone[p]
            self.connect(p, handle, off[2 * self.rank], off[2 * self.rank + 1], off[2 * self.nparts])
        dist.barrier(group=group)

    @staticmethod
    def wire_local(shards):
        """Single-process wiring of all shards of a partition (tests / one GPU): id lists are handed over directly, the
        exchange itself then goes through pack / unpack (exchange_local) instead of peer memory."""
        for s in shards:
            for p in range(s.nparts):
                if p != s.rank and s.recv_count(p) > 0:
                    shards[p].set_send_ids(s.rank, s.recv_ids(p))
        for s in shards:
            s.finalize()

    # ---- application ----
    @staticmethod
    def _terms(which, coef):
        import ctypes as C
        return (C.c_int32 * len(which))(*which), (C.c_double * len(coef))(*coef)

    def spmv_device(self, op, which, coef, x_ptr, y_ptr):
        """y[0:n_owned] = sum_i coef[i] D[which[i]] [x ; halo]: ONE launch, halo exchange fused (NVLink peer stores)"""
        w, c = self._terms(which, coef)
        self.ctx._check(self.ctx._L.rbffd_shard_spmv_device(self._h, op._h, len(which), w, c, x_ptr, y_ptr))

    def spmv_stage_device(self, op, which, coef, x_ptr, a, u_ptr, b, dt, out_ptr):
        """out[0:n_owned] = a*u + b*(x + dt * sum_i coef[i] D[which[i]] [x ; halo]): halo exchange + product + SSP-RK stage, ONE launch"""
        w, c = self._terms(which, coef)
        self.ctx._check(self.ctx._L.rbffd_shard_spmv_stage_device(self._h, op._h, len(which), w, c, x_ptr, a, u_ptr, b, dt, out_ptr))

    def spmv_local_device(self, op, which, coef, x_ptr, y_ptr):
        w, c = self._terms(which, coef)
        self.ctx._check(self.ctx._L.rbffd_shard_spmv_local_device(self._h, op._h, len(which), w, c, x_ptr, y_ptr))

    def spmv_t_device(self, op, which, v_ptr, y_ptr, alpha=1.0, beta=0.0):
        """y[0:n_owned] = alpha D[which]' v + beta y with the reverse halo exchange (E' * v of adv_diff_test.jl:151)"""
        self.ctx._check(self.ctx._L.rbffd_shard_spmv_t_device(self._h, op._h, which, alpha, v_ptr, beta, y_ptr))

    def spmv_t_local_device(self, op, which, v_ptr, y_ptr, alpha=1.0, beta=0.0):
        self.ctx._check(self.ctx._L.rbffd_shard_spmv_t_local_device(self._h, op._h, which, alpha, v_ptr, beta, y_ptr))

    def pack_device(self, peer, x_ptr, buf_ptr):
        self.ctx._check(self.ctx._L.rbffd_shard_pack_device(self._h, peer, x_ptr, buf_ptr))

    def unpack_device(self, peer, buf_ptr):
        self.ctx._check(self.ctx._L.rbffd_shard_unpack_device(self._h, peer, buf_ptr))

    def tpack_device(self, peer, buf_ptr):
        self.ctx._check(self.ctx._L.rbffd_shard_tpack_device(self._h, peer, buf_ptr))

    def tunpack_add_device(self, peer, buf_ptr, y_ptr):
        self.ctx._check(self.ctx._L.rbffd_shard_tunpack_add_device(self._h, peer, buf_ptr, y_ptr))

    def exchange_collective(self, x, group=None):
        """Forward halo exchange over torch.distributed point-to-point ops (NCCL send/recv over NVLink, gloo, ...) instead of
        peer memory: pack -> batched isend/irecv -> unpack into the inbox.  Follow with spmv_local_device.  x: torch tensor of
        the owned values.  The transport-agnostic alternative to the fused launch (e.g. no CUDA IPC between the ranks)."""
        import torch
        import torch.distributed as dist
        ops, recv, keep = [], [], []
        for p in range(self.nparts):
            if p == self.rank:
                continue
            ns, nr = self.send_count(p), self.recv_count(p)
            if ns:
                sb = torch.empty(ns, dtype=torch.float64, device=x.device)
                self.pack_device(p, x.data_ptr(), sb.data_ptr())
                keep.append(sb)
                ops.append(dist.P2POp(dist.isend, sb, p if group is None else dist.get_global_rank(group, p), group))
            if nr:
                rb_ = torch.empty(nr, dtype=torch.float64, device=x.device)
                recv.append((p, rb_))
                ops.append(dist.P2POp(dist.irecv, rb_, p if group is None else dist.get_global_rank(group, p), group))
        if ops:
            self.ctx.synchronize()                         # the packed values are complete before the transport reads them
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p, rb_ in recv:
            self.unpack_device(p, rb_.data_ptr())
        self._bufs = keep + [r for _, r in recv]

    @staticmethod
    def exchange_local(shards, xs):
        """forward halo exchange of a single-process partition: xs[r] = owned values of shard r (torch, on the device)"""
        import torch
        for s in shards:
            for p in range(s.nparts):
                cnt = s.recv_count(p)
                if p == s.rank or cnt == 0:
                    continue
                buf = torch.empty(cnt, dtype=torch.float64, device=xs[p].device)
                shards[p].pack_device(s.rank, xs[p].data_ptr(), buf.data_ptr())
                s.unpack_device(p, buf.data_ptr())
                s._bufs.append(buf)
        for s in shards:
            s.ctx.synchronize()
            s._bufs.clear()

    @staticmethod
    def exchange_t_local(shards, ys):
        """reverse exchange of a single-process partition after spmv_t_local_device on every shard: the halo-column sums go
        back to their owners and are added in ascending peer order"""
        import torch
        for s in shards:                       # s = owner receiving
            for p in range(s.nparts):          # p = peer that holds s's nodes in its halo
                cnt = s.send_count(p)
                if p == s.rank or cnt == 0:
                    continue
                buf = torch.empty(cnt, dtype=torch.float64, device=ys[s.rank].device)
                shards[p].tpack_device(s.rank, buf.data_ptr())
                s.tunpack_add_device(p, buf.data_ptr(), ys[s.rank].data_ptr())
                s._bufs.append(buf)
        for s in shards:
            s.ctx.synchronize()
            s._bufs.clear()
