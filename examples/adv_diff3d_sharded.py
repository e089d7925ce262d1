#!/usr/bin/env python
"""BASELINE.json configs[4] at any size: 3-D advection-diffusion time stepping with sharded weight generation and
halo-exchange SpMV, one process per GPU.

    python examples/adv_diff3d_sharded.py --g 64 --steps 20                       # one GPU, 64^3 nodes
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/adv_diff3d_sharded.py --g 58 --steps 20                          # 8 GPUs, (58*2)^3 = 1.56 M nodes ... --g 232: 100 M

What runs (all on the device, nothing of size O(nodes) crosses PCIe):
  * the lattice is cut into spatial blocks (1, 2x1x1, 2x2x1, 2x2x2 for 1/2/4/8 GPUs); every rank generates its block plus a
    candidate margin from the closed-form node generator, finds the stencils of its owned nodes (exact kNN, n = 60, ties by
    global id), proves the margin wide enough and keeps the stencil closure as its halo (rbffd_shard_create_device)
    -> no communication;
  * one fused weight solve per owned node writes the Laplacian and the three first derivatives (PHS r^7 + degree-3
    polynomials; the reference calls: generate_operator.jl:29-190 in 3-D);
  * u_t = alpha Lap u - a . grad u  (the interior line of cons_sys, examples/adv_diff_test.jl:151-152, in 3-D): the four
    value arrays are combined once into one matrix (constant coefficients; --no-combine applies the four operators in
    every stage), and every SSP-RK3 stage is ONE launch (rbffd_shard_spmv_stage_device): the first CTAs store this rank's
    boundary values into the neighbours' inboxes over NVLink, the interior rows run meanwhile, the boundary rows wait
    for the neighbours' flags, and the stage combination a*u + b*(v + dt*L v) is the epilogue of the product;
  * the whole step is captured into a CUDA graph (--graph) on any number of GPUs: the exchange epoch lives in device memory;
  * nodes within `--bw` of the cube's faces carry the exact solution (a Gaussian pulse advected by a and spread by alpha),
    which is also the error reference at the end.
Prints ONE JSON line on rank 0."""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbffd_b200 as rb  # noqa: E402

P, N_ST, DEG = 7, 60, 3
OPS = ["Lap", "Dx", "Dy", "Dz"]


def exact(X, t, alpha, a, x0, s0):
    """Gaussian pulse: solves u_t + a.grad u = alpha Lap u in free space."""
    s2 = s0 * s0 + 2.0 * alpha * t
    d2 = sum((X[:, c] - (x0[c] + a[c] * t)) ** 2 for c in range(3))
    return (s0 * s0 / s2) ** 1.5 * torch.exp(-d2 / (2.0 * s2))


BLOCKS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def device_view(ptr, shape, dtype, dev):
    """zero-copy torch view of library-owned device memory"""
    import numpy as np

    class _Raw:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": np.dtype(dtype).str, "data": (int(ptr), False), "version": 3, "strides": None}
    return torch.as_tensor(_Raw(), device=dev)


def run(g, steps, alpha=2e-3, a=(0.3, 0.2, 0.1), bw=None, cfl=0.08, combine=True, graph=False, blocks=None):
    import torch.distributed as dist
    from rbffd_b200 import sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    dim, n = 3, N_ST
    G = int(round(g * world ** (1.0 / 3.0)))
    blocks = tuple(blocks) if blocks else BLOCKS[world]
    ctx = rb.Context(lr, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t_gen = time.perf_counter()
    shard = sharding.Shard.lattice_block(ctx, dim, G, 0, blocks, rank, n)
    M = shard.n_owned
    op4 = shard.generate(P, DEG, OPS, dim)
    if combine:
        # constant coefficients: alpha*Lap - ax*Dx - ay*Dy - az*Dz becomes ONE matrix (the reference rebuilds this sparse
        # sum in every cons_sys call, adv_diff_test.jl:151-152), every stage is then a single-matrix product
        vc = torch.empty((1, M, n), dtype=torch.float64, device=dev)
        op4.combine_device([0, 1, 2, 3], [alpha, -a[0], -a[1], -a[2]], vc.data_ptr())
        ci = op4._keep[0]
        op4.close()
        op4._keep = None
        op = ctx.operator_from_device(M, M + shard.n_halo, n, 1, ci.data_ptr(), vc.data_ptr())
        op._keep = (ci, vc)
        which, coef = [0], [1.0]
    else:
        op, which, coef = op4, [0, 1, 2, 3], [alpha, -a[0], -a[1], -a[2]]
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    t_wire = time.perf_counter()
    if world > 1:
        shard.wire()
    else:
        shard.finalize()
    t_wire = time.perf_counter() - t_wire

    own = device_view(shard.X_local_ptr, (M + shard.n_halo, dim), "f8", dev)[:M]       # owned nodes in local order [interior | boundary]
    h = 1.0 / G
    bw = 2.5 * h if bw is None else bw
    inner = ((own > bw) & (own < 1.0 - bw)).all(dim=1)
    bidx = torch.nonzero(~inner).squeeze(1)          # Dirichlet layer
    Xb = own[bidx]
    x0, s0 = (0.35, 0.4, 0.45), max(0.08, 3.0 * h)
    dt = cfl * h * h / alpha
    u = exact(own, 0.0, alpha, a, x0, s0)
    v1, v2 = torch.empty_like(u), torch.empty_like(u)
    t_dev = torch.zeros((), dtype=torch.float64, device=dev)     # the time lives on the device: the step is replayable
    Pt = lambda t: t.data_ptr()

    def dirichlet(v, t):                             # boundary data: exact solution at the stage time
        v.index_copy_(0, bidx, exact(Xb, t, alpha, a, x0, s0))

    def rk3_step():                                  # SSP-RK3 (Shu-Osher): every stage = ONE launch (exchange + product + combination)
        shard.spmv_stage_device(op, which, coef, Pt(u), 0.0, Pt(u), 1.0, dt, Pt(v1))               # v1 = u + dt L(u)
        dirichlet(v1, t_dev + dt)
        shard.spmv_stage_device(op, which, coef, Pt(v1), 0.75, Pt(u), 0.25, dt, Pt(v2))            # v2 = 3/4 u + 1/4 (v1 + dt L(v1))
        dirichlet(v2, t_dev + 0.5 * dt)
        shard.spmv_stage_device(op, which, coef, Pt(v2), 1.0 / 3.0, Pt(u), 2.0 / 3.0, dt, Pt(v1))  # u' = 1/3 u + 2/3 (v2 + dt L(v2))
        dirichlet(v1, t_dev + dt)
        u.copy_(v1)
        t_dev.add_(dt)

    replay, todo = rk3_step, steps
    if graph and steps > 1:
        # a fixed sequence of launches with static buffers, the exchange epoch in device memory -> captured ONCE into a CUDA
        # graph (the context launches on the capture stream) and replayed, on any number of GPUs
        rk3_step()
        todo -= 1
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        cg = torch.cuda.CUDAGraph()
        main_stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.graph(cg):
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            rk3_step()
        ctx.set_stream(main_stream)
        replay = cg.replay
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(todo):
        replay()
    ev1.record()
    torch.cuda.synchronize()
    t_run = ev0.elapsed_time(ev1) * 1e-3
    shard.check()                                    # a halo exchange that timed out waiting for a peer is an error
    if world > 1:
        dist.barrier()
    t = steps * dt
    steps_timed = max(todo, 1)
    ue = exact(own, t, alpha, a, x0, s0)
    acc = torch.stack([((u - ue) ** 2).sum(), (ue ** 2).sum(), u.sum(), (u * (own[:, 0] + 2 * own[:, 1] + 3 * own[:, 2])).sum()])
    tt = torch.tensor([t_gen, t_run, t_wire], dtype=torch.float64, device=dev)
    cnt = torch.tensor([M, shard.n_interior, shard.n_halo], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(acc)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
    err = math.sqrt(float(acc[0]) / float(acc[1]))
    ms_step = float(tt[1]) / steps_timed * 1e3
    nmat = len(which)
    bytes_per_row = 8 * n * nmat + 4 * n + 8 + 3 * 8          # values + indices + gathered x (counted once) + stage epilogue (u, x, out)
    out = {"example": "adv_diff3d_sharded", "n_gpus": world, "blocks": list(blocks), "global_nodes": G ** 3, "nodes_per_gpu_max": int(cnt[0]),
           "halo_nodes_max": int(cnt[2]), "interior_rows_max": int(cnt[1]), "n": n, "p": P, "polydeg": DEG,
           "steps": steps, "dt": dt, "t_end": t, "rel_l2_error_vs_exact": err, "checksum": [float(acc[2]), float(acc[3])],
           "generation_s": float(tt[0]), "stencils_per_s": G ** 3 / float(tt[0]), "halo_wiring_s": float(tt[2]),
           "ms_per_step": ms_step, "rhs_evaluations_per_s": 3 * steps_timed / float(tt[1]),
           "spmv_halo_gbs_per_gpu": 3 * int(cnt[0]) * bytes_per_row / (ms_step * 1e-3) * 1e-9,
           "cuda_graph": bool(graph and steps > 1), "operators_per_stage": nmat, "launches_per_stage": 1,
           "halo": "fused into the SpMV launch: NVLink peer stores into CUDA-IPC inboxes, device-side epoch" if world > 1 else "none"}
    uh = u.clone()
    op.close()
    shard.close()
    return out, uh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--g", type=int, default=48, help="lattice size per GPU: g^3 nodes per rank")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-combine", action="store_true", help="apply the four operators in every stage (fused multi-operator SpMV) "
                    "instead of pre-combining them into one matrix")
    ap.add_argument("--graph", action="store_true", help="capture the SSP-RK3 step into a CUDA graph and replay it (any number of GPUs)")
    args = ap.parse_args()
    out, _ = run(args.g, args.steps, combine=not args.no_combine, graph=args.graph)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out))
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
