"""General spatial-block shards (csrc/shard.cu) on ONE GPU: all shards of a partition live in one process, the halo values
travel through pack / unpack (the transport-agnostic form of the exchange).  Bars: stencils, pattern and weights of every shard
are BIT-IDENTICAL to the single-GPU operator on the same node set (ties by global id), the sharded D*u equals the single-GPU
product bit for bit (same row arithmetic), the sharded E'*v to 1e-13 of |A|'|v| (the column sums run in another order).
The real NVLink peer-memory exchange is covered by tests/mgpu_shard_check.py (2 GPUs)."""
import os

import numpy as np
import pytest

import rbffd_b200 as rb
from rbffd_b200 import sharding

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    import torch
    c = rb.Context(0)
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    return c


def _check_partition(ctx, X, part, nparts, p, n, polydeg, ops, shards=None):
    import torch
    dim = X.shape[1]
    N = len(X)
    colind, vals = rb.generate_raw(X, None, p, n, polydeg, ops, ctx=ctx)
    gop = rb.Operator.from_host(ctx, colind, vals, N)
    rng = np.random.default_rng(5)
    u = rng.standard_normal(N)
    v = rng.standard_normal(N)
    own = shards is not None
    if not own:
        shards = [sharding.Shard.from_host(ctx, X, part, nparts, r, n) for r in range(nparts)]
    sharding.Shard.wire_local(shards)
    lops, xs, ys, gids = [], [], [], []
    seen = np.zeros(N, int)
    for s in shards:
        g = s.global_ids()
        gids.append(g)
        assert s.n_owned + s.n_halo == len(g) and 0 <= s.n_interior <= s.n_owned
        if part is not None:
            assert np.array_equal(np.sort(g[:s.n_owned]), np.flatnonzero(part == s.rank))
        seen[g[:s.n_owned]] += 1
        op = s.generate(p, polydeg, ops, dim)
        lc, lv = op.to_host()
        # rows in local order [interior | boundary]; columns local -> global
        assert np.array_equal(g[lc], colind[g[:s.n_owned]]), "sharded stencils differ from the single-GPU stencils"
        assert np.array_equal(lv, vals[:, g[:s.n_owned]]), "sharded weights differ from the single-GPU weights"
        # interior rows reference no halo column, every boundary row references at least one
        assert lc[:s.n_interior].max(initial=-1) < s.n_owned
        if s.n_owned > s.n_interior:
            assert (lc[s.n_interior:].max(1) >= s.n_owned).all()
        # the halo is exactly the stencil closure, grouped by owner, ascending global id inside every group
        assert np.array_equal(np.unique(lc[lc >= s.n_owned]), np.arange(s.n_owned, s.n_owned + s.n_halo))
        if part is not None and s.n_halo:
            ho = part[g[s.n_owned:]]
            assert (np.diff(ho) >= 0).all() and not (ho == s.rank).any()
            for o in np.unique(ho):
                assert (np.diff(g[s.n_owned:][ho == o]) > 0).all()
        lops.append(op)
        xs.append(torch.from_numpy(u[g[:s.n_owned]]).cuda())
        ys.append(torch.empty(s.n_owned, dtype=torch.float64, device="cuda"))
    assert (seen == 1).all()
    # D * u : exchange, then the sharded product (no exchange inside)
    sharding.Shard.exchange_local(shards, xs)
    which, coef = list(range(min(len(ops), 3))), [0.7, -1.3, 0.4][:min(len(ops), 3)]
    xg = torch.from_numpy(u).cuda()
    yg = torch.empty(N, dtype=torch.float64, device="cuda")
    gop.spmv_multi_device(which, coef, xg.data_ptr(), yg.data_ptr())
    ctx.synchronize()
    for s, op, x, y, g in zip(shards, lops, xs, ys, gids):
        s.spmv_local_device(op, which, coef, x.data_ptr(), y.data_ptr())
        ctx.synchronize()
        assert np.array_equal(y.cpu().numpy(), yg.cpu().numpy()[g[:s.n_owned]]), "sharded D*u differs from the single-GPU product"
    # E' * v with the reverse exchange
    ref = gop.spmv_t(0, v, alpha=1.5)
    bound = rb.Operator.from_host(ctx, colind, np.abs(vals[:1]), N).spmv_t(0, np.abs(v), alpha=1.5)
    vs = [torch.from_numpy(v[g[:s.n_owned]]).cuda() for s, g in zip(shards, gids)]
    for s, op, vv, y in zip(shards, lops, vs, ys):
        s.spmv_t_local_device(op, 0, vv.data_ptr(), y.data_ptr(), alpha=1.5)
    sharding.Shard.exchange_t_local(shards, ys)
    for s, y, g in zip(shards, ys, gids):
        got = y.cpu().numpy()
        assert np.all(np.abs(got - ref[g[:s.n_owned]]) <= 1e-13 * bound[g[:s.n_owned]] + 1e-300), "sharded E'*v differs"
    for s in shards:
        s.check()
    for op in lops:
        op.close()
    if not own:
        for s in shards:
            s.close()


@pytest.mark.parametrize("nparts,blocks", [(2, None), (3, None), (4, (2, 2)), (5, (5, 1))])
def test_shards_of_the_tominec_node_set(ctx, tominec, nparts, blocks):
    """the reference's own scattered node set (test/data/x_nodes_fitted.csv), p = 3, n = 20, degree 3 (poisson_test.jl:51-53)"""
    X = tominec["X"]
    part = sharding.plan(X, nparts, blocks)
    _check_partition(ctx, X, part, nparts, 3, 20, 3, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"])


def test_shards_of_the_mesh_node_set(ctx):
    """BASELINE configs[0]: the node set of examples/rect_0_10.cgns (centroids + boundary + ghost nodes), p = 5, n = 42, degree 5"""
    Y, P, iin, ibc, ig, cells, nrm, tan = rb.mesh.processmesh(os.path.join(ROOT, "tests", "golden", "rect_0_10.cgns"), ["left", "right", "top", "bottom"])
    X = np.ascontiguousarray(Y, np.float64)
    part = sharding.plan(X, 3)
    _check_partition(ctx, X, part, 3, 5, 42, 5, ["Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)])


def test_halton_like_scattered_set_3d(ctx):
    """3-D scattered nodes that are NOT a lattice (van der Corput / Halton bases 2, 3, 5), 2 x 2 x 2 blocks, n = 60, r^7, degree 3"""
    def vdc(i, b):
        f, r = 1.0, np.zeros(len(i))
        i = i.copy()
        while i.any():
            f /= b
            r += f * (i % b)
            i //= b
        return r
    i = np.arange(1, 9001)
    X = np.column_stack([vdc(i, 2), vdc(i, 3), vdc(i, 5)])
    part = sharding.plan(X, 8, blocks=(2, 2, 2))
    _check_partition(ctx, X, part, 8, 7, 60, 3, ["Lap", "Dx", "Dy", "Dz"])


def test_lattice_blocks_generated_on_the_device(ctx):
    """BASELINE configs[3]/[4] shape: every rank generates its own 2x2x2 block of the closed-form lattice + candidate margin on
    the device (no node ever crosses the host); the union reproduces the single-GPU operator of the full lattice."""
    g, n = 22, 60
    X = rb.nodes.jittered_lattice(3, g, seed=0)
    shards = [sharding.Shard.lattice_block(ctx, 3, g, 0, (2, 2, 2), r, n) for r in range(8)]
    c = np.minimum((X * g).astype(int), g - 1)
    part = (((c[:, 2] * 2) // g) + 2 * ((c[:, 1] * 2) // g) + 4 * ((c[:, 0] * 2) // g)).astype(np.int32)
    for s in shards:
        gid = s.global_ids()
        assert np.array_equal(np.sort(gid[:s.n_owned]), np.flatnonzero(part == s.rank))
    _check_partition(ctx, X, part, 8, 7, n, 3, ["Lap", "Dx", "Dy", "Dz"], shards=shards)
    for s in shards:
        s.close()


def test_thin_halo_is_refused(ctx):
    """a candidate set that does not enclose the stencils must be refused (RBFFD_ERR_HALO), never silently accepted"""
    import torch
    g, n = 40, 30
    with pytest.raises(rb.RbffdError) as e:
        # margin 0 and no retry: build the box by hand
        import ctypes as C
        dev = torch.device("cuda", 0)
        lo, hi = (C.c_int64 * 3)(0, 0, 0), (C.c_int64 * 3)(g // 2, g, 0)
        nc = (g // 2) * g
        Xc = torch.empty(nc * 2, dtype=torch.float64, device=dev)
        gid = torch.empty(nc, dtype=torch.int64, device=dev)
        own = torch.empty(nc, dtype=torch.int32, device=dev)
        ctx._check(ctx._L.rbffd_jittered_lattice_box_device(ctx._h, 2, g, 0, lo, hi, (C.c_int32 * 3)(2, 1, 1), Xc.data_ptr(), gid.data_ptr(), own.data_ptr()))
        sharding.Shard.from_device(ctx, 2, n, Xc.data_ptr(), gid.data_ptr(), own.data_ptr(), nc, [-np.inf, -np.inf], [(g // 2 + 0.25) / g, np.inf], 0, 2)
    assert e.value.code == rb._lib.ERR_HALO


def test_fused_halo_exchange_two_gpus():
    """the real thing: one process per GPU, CUDA-IPC inboxes, halo exchange fused into the SpMV launch (tests/mgpu_shard_check.py)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29581", os.path.join(ROOT, "tests", "mgpu_shard_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
