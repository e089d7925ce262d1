"""Multi-GPU check of the general spatial-block shards (csrc/shard.cu), launched with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tests/mgpu_shard_check.py
Every rank builds its shard, wires the CUDA-IPC inboxes, generates its operator rows without communication and applies them with
the halo exchange FUSED into the SpMV launch (NVLink peer stores + epoch flags).  Bars: stencils / weights bit-identical to the
single-GPU operator of the full node set, D*u bit-identical over many epochs (eager and CUDA-graph replay), E'*v to 1e-13 |A|'|v|.
Prints one line per case and a final OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbffd_b200 as rb  # noqa: E402
from rbffd_b200 import sharding  # noqa: E402


def check(ctx, rank, world, name, X, shard, p, polydeg, ops, dim, epochs=6):
    n = shard.n
    N = len(X)
    shard.wire()
    op = shard.generate(p, polydeg, ops, dim)
    g = shard.global_ids()
    colind, vals = rb.generate_raw(X, None, p, n, polydeg, ops, ctx=ctx)          # single-GPU truth, computed by every rank
    lc, lv = op.to_host()
    assert np.array_equal(g[lc], colind[g[:shard.n_owned]]), f"{name}: stencils differ"
    assert np.array_equal(lv, vals[:, g[:shard.n_owned]]), f"{name}: weights differ"
    gop = rb.Operator.from_host(ctx, colind, vals, N)
    which, coef = list(range(min(3, len(ops)))), [0.7, -1.3, 0.4][:min(3, len(ops))]
    rng = np.random.default_rng(11)
    x = torch.empty(shard.n_owned, dtype=torch.float64, device="cuda")
    y = torch.empty(shard.n_owned, dtype=torch.float64, device="cuda")
    yg = torch.empty(N, dtype=torch.float64, device="cuda")
    own = torch.from_numpy(g[:shard.n_owned]).cuda()
    fields = [rng.standard_normal(N) for _ in range(epochs)]
    for u in fields:                                   # back-to-back epochs, no host synchronisation in between
        ug = torch.from_numpy(u).cuda()
        x.copy_(ug[own])
        shard.spmv_device(op, which, coef, x.data_ptr(), y.data_ptr())
        gop.spmv_multi_device(which, coef, ug.data_ptr(), yg.data_ptr())
        assert torch.equal(y, yg[own]), f"{name}: sharded D*u differs from the single-GPU product"
    # the same product over the torch.distributed transport (NCCL send/recv): pack -> isend/irecv -> unpack -> local product
    ug = torch.from_numpy(fields[-1]).cuda()
    x.copy_(ug[own])
    shard.exchange_collective(x)
    shard.spmv_local_device(op, which, coef, x.data_ptr(), y.data_ptr())
    gop.spmv_multi_device(which, coef, ug.data_ptr(), yg.data_ptr())
    assert torch.equal(y, yg[own]), f"{name}: sharded D*u over the collective transport differs"
    # CUDA-graph replay: the epoch lives in device memory, so the captured launch is replayable
    ug = torch.from_numpy(fields[0]).cuda()
    x.copy_(ug[own])
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    cs = torch.cuda.Stream()
    with torch.cuda.stream(cs):
        ctx.set_stream(cs.cuda_stream)
        with torch.cuda.graph(graph, stream=cs):
            shard.spmv_device(op, which, coef, x.data_ptr(), y.data_ptr())
            x.copy_(y)                                  # two dependent applications per replay: y2 = D (D u)
            shard.spmv_device(op, which, coef, x.data_ptr(), y.data_ptr())
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for rep in range(3):
        x.copy_(ug[own])
        graph.replay()
    torch.cuda.synchronize()
    y1 = torch.empty(N, dtype=torch.float64, device="cuda")
    gop.spmv_multi_device(which, coef, ug.data_ptr(), y1.data_ptr())
    gop.spmv_multi_device(which, coef, y1.data_ptr(), yg.data_ptr())
    ctx.synchronize()
    assert torch.equal(y, yg[own]), f"{name}: graph replay of the sharded product differs"
    # E' * v, twice (epochs of the reverse exchange)
    for rep in range(2):
        v = rng.standard_normal(N)
        ref = gop.spmv_t(0, v, alpha=1.5)
        bound = rb.Operator.from_host(ctx, colind, np.abs(vals[:1]), N).spmv_t(0, np.abs(v), alpha=1.5)
        vv = torch.from_numpy(v[g[:shard.n_owned]]).cuda()
        shard.spmv_t_device(op, 0, vv.data_ptr(), y.data_ptr(), alpha=1.5)
        ctx.synchronize()
        got = y.cpu().numpy()
        assert np.all(np.abs(got - ref[g[:shard.n_owned]]) <= 1e-13 * bound[g[:shard.n_owned]] + 1e-300), f"{name}: sharded E'*v differs"
    shard.check()                                      # no exchange timed out
    dist.barrier()
    if rank == 0:
        print(f"{name}: world {world}, N {N}, rank 0 owns {shard.n_owned} ({shard.n_interior} interior) + {shard.n_halo} halo: bit-identical", flush=True)
    op.close()
    gop.close()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = rb.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    # 1. the reference's scattered node set (test/data/x_nodes_fitted.csv)
    X = np.load(os.path.join(ROOT, "tests", "golden", "tominec_fitted.npz"))["X"]
    part = sharding.plan(X, world)
    s = sharding.Shard.from_host(ctx, X, part, world, rank, 20)
    check(ctx, rank, world, "tominec p3 n20", X, s, 3, 3, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"], 2)
    s.close()
    # 2. BASELINE configs[0] node set (mesh centroids + boundary + ghost nodes)
    Y = rb.mesh.processmesh(os.path.join(ROOT, "tests", "golden", "rect_0_10.cgns"), ["left", "right", "top", "bottom"])[0]
    X = np.ascontiguousarray(Y, np.float64)
    part = sharding.plan(X, world)
    s = sharding.Shard.from_host(ctx, X, part, world, rank, 42)
    check(ctx, rank, world, "rect_0_10 p5 n42", X, s, 5, 5, ["Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)], 2)
    s.close()
    # 3. 3-D lattice blocks generated on the device (configs[3]/[4] shape)
    g = int(os.environ.get("SHARD_G", "26"))
    blocks = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    X = rb.nodes.jittered_lattice(3, g, seed=0)
    s = sharding.Shard.lattice_block(ctx, 3, g, 0, blocks, rank, 60)
    check(ctx, rank, world, f"lattice {g}^3 blocks {blocks} p7 n60", X, s, 7, 3, ["Lap", "Dx", "Dy", "Dz"], 3)
    s.close()
    dist.barrier()
    if rank == 0:
        print("OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
