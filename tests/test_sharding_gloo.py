"""N > 1 host logic on CPU: slab partition invariants and the halo exchange under gloo, world_size 2 and 3."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rbffd_b200 as rb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_covers_lattice():
    for dim, g, world, halo in ((2, 37, 4, 5), (3, 11, 2, 3), (2, 16, 8, 2), (2, 10, 1, 4)):
        shards = [rb.SlabShard(r, world, dim, g, halo) for r in range(world)]
        assert sum(s.n_owned for s in shards) == g**dim
        assert shards[0].n_lo == 0 and shards[-1].n_hi == 0
        for a, b in zip(shards[:-1], shards[1:]):
            assert a.row1 == b.row0
            assert a.first_owned_id + a.n_owned == b.first_owned_id
        for s in shards:
            X = rb.nodes.jittered_lattice(dim, g, 0, s.first_local_id, s.n_local)
            full = rb.nodes.jittered_lattice(dim, g, 0)
            assert np.array_equal(X, full[s.first_local_id:s.first_local_id + s.n_local])
            below, above = s.missing_edges()
            if below is not None:
                assert full[:s.first_local_id, -1].max() <= below
            if above is not None:
                assert full[s.first_local_id + s.n_local:, -1].min() >= above


def test_boundary_row_ranges():
    for world, rank in ((4, 0), (4, 2), (4, 3), (1, 0)):
        s = rb.SlabShard(rank, world, 2, 40, 3)
        (a0, a1), (b0, b1), (c0, c1) = rb.boundary_row_ranges(s)
        assert a0 == 0 and a1 == b0 and b1 == c0 and c1 == s.n_owned
        assert a1 - a0 == s.n_lo and c1 - c0 == s.n_hi


def test_halo_sufficiency_check(oracle):
    dim, g, world = 2, 40, 2
    full = rb.nodes.jittered_lattice(dim, g, 0)
    ref = oracle.knn(full, full, 30)[0]
    for halo, expect in ((1, False), (8, True)):
        s = rb.SlabShard(0, world, dim, g, halo)
        Xl = full[s.first_local_id:s.first_local_id + s.n_local]
        own = Xl[s.n_lo:s.n_lo + s.n_owned]
        idx, d2 = oracle.knn(Xl, own, 30)
        ok = s.halo_is_sufficient(own[:, -1], d2[:, -1])
        assert ok == expect
        if ok:      # sufficient halo  =>  local stencils are the global stencils (shifted ids): zero communication
            assert np.array_equal(idx + s.first_local_id, ref[s.first_owned_id:s.first_owned_id + s.n_owned])


def _worker(rank, world, port, dim, g, halo):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import rbffd_b200 as rbw
    s = rbw.SlabShard(rank, world, dim, g, halo)
    gid = torch.arange(s.first_local_id, s.first_local_id + s.n_local, dtype=torch.float64)
    u = torch.full((s.n_local,), -1.0, dtype=torch.float64)
    u[s.n_lo:s.n_lo + s.n_owned] = gid[s.n_lo:s.n_lo + s.n_owned] * 3.0 + 1.0      # owned values: f(global id)
    rbw.exchange_halo(u, s)
    ok = torch.equal(u, gid * 3.0 + 1.0)
    u[:s.n_lo] = -1.0
    u[s.n_lo + s.n_owned:] = -1.0
    rbw.exchange_halo(u, s, async_op=True).wait()
    ok = ok and torch.equal(u, gid * 3.0 + 1.0)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if flag.item() != 1:
        raise SystemExit(3)


@pytest.mark.parametrize("world,dim,g,halo", [(2, 2, 24, 4), (3, 3, 9, 2)])
def test_halo_exchange_gloo(world, dim, g, halo):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, dim, g, halo), nprocs=world, join=True)


def _sharded_apply_worker(rank, world, port, dim, g, halo, n, p, deg, steps):
    """The N > 1 application path of examples/adv_diff3d_sharded.py on CPU: slab-local operator rows (oracle weights on the local
    node set, local column ids), halo exchange of the field, boundary rows after the exchange -- must reproduce the global
    operator applied to the global field, step after step."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import rbffd_b200 as rbw
    from oracle import oracle as orc
    s = rbw.SlabShard(rank, world, dim, g, halo)
    full = rbw.nodes.jittered_lattice(dim, g, 0)
    Xl = full[s.first_local_id:s.first_local_id + s.n_local]
    own = Xl[s.n_lo:s.n_lo + s.n_owned]
    ops = ["Lap", "Dx"]
    ci, va = orc.generate_operator(Xl, own, p, n, deg, ops=ops)           # rows = owned nodes, columns = local ids
    gci, gva = orc.generate_operator(full, full, p, n, deg, ops=ops)      # the global operator (reference)
    ok = np.array_equal(ci + s.first_local_id, gci[s.first_owned_id:s.first_owned_id + s.n_owned])   # zero-communication generation
    coef = (0.01, -0.3)
    u_glob = np.sin(3 * full[:, 0]) * np.cos(2 * full[:, -1])
    u = torch.zeros(s.n_local, dtype=torch.float64)
    u[s.n_lo:s.n_lo + s.n_owned] = torch.from_numpy(u_glob[s.first_owned_id:s.first_owned_id + s.n_owned])
    (l0, l1), (i0, i1), (h0, h1) = rbw.boundary_row_ranges(s)
    dt = 1e-3
    for _ in range(steps):
        work = rbw.exchange_halo(u, s, async_op=True)
        du = np.zeros(s.n_owned)
        un = u.numpy()
        if i1 > i0:                                                       # interior rows never touch a halo column ...
            assert ci[i0:i1].min() >= s.n_lo and ci[i0:i1].max() < s.n_lo + s.n_owned
            du[i0:i1] = sum(c * orc.spmv(ci[i0:i1], v[i0:i1], np.where(np.arange(s.n_local) < s.n_lo, np.nan, np.where(np.arange(s.n_local) >= s.n_lo + s.n_owned, np.nan, un)))
                            for c, v in zip(coef, va))                    # ... (halo entries poisoned with NaN to prove it)
        work.wait()
        for (r0, r1) in ((l0, l1), (h0, h1)):
            if r1 > r0:
                du[r0:r1] = sum(c * orc.spmv(ci[r0:r1], v[r0:r1], un) for c, v in zip(coef, va))
        dg = sum(c * orc.spmv(gci, v, u_glob) for c, v in zip(coef, gva))
        ok = ok and np.array_equal(du, dg[s.first_owned_id:s.first_owned_id + s.n_owned])
        u[s.n_lo:s.n_lo + s.n_owned] += dt * torch.from_numpy(du)
        u_glob = u_glob + dt * dg
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if flag.item() != 1:
        raise SystemExit(3)


@pytest.mark.parametrize("world,dim,g,halo,n,p,deg", [(2, 2, 24, 5, 12, 3, 1), (2, 3, 10, 4, 20, 3, 1)])
def test_sharded_operator_application_gloo(world, dim, g, halo, n, p, deg):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_sharded_apply_worker, args=(world, port, dim, g, halo, n, p, deg, 3), nprocs=world, join=True)
