#!/usr/bin/env python
"""BASELINE.json configs[4] at any size: 3-D advection-diffusion time stepping with sharded weight generation and
halo-exchange SpMV, one process per GPU.

    python examples/adv_diff3d_sharded.py --g 64 --steps 20                       # one GPU, 64^3 nodes
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/adv_diff3d_sharded.py --g 58 --steps 20                          # 8 GPUs, (58*2)^3 = 1.56 M nodes ... --g 232: 100 M

What runs (all on the device, nothing of size O(nodes) crosses PCIe):
  * every rank generates its slab [halo | owned | halo] of the jittered lattice from the closed-form node generator,
    finds the stencils of its owned nodes (exact kNN, n = 60) and proves the halo wide enough -> no communication;
  * one fused weight kernel launch per row range (low boundary / interior / high boundary) writes the Laplacian and the
    three first derivatives (PHS r^7 + degree-3 polynomials; the reference calls: generate_operator.jl:29-190 in 3-D);
  * u_t = alpha Lap u - a . grad u  (the interior line of cons_sys, examples/adv_diff_test.jl:151-152, in 3-D): the four
    value arrays are combined once into one matrix (constant coefficients), so every stage is ONE single-matrix SpMV
    (--no-combine: one fused four-operator SpMV per stage over the shared pattern); the interior rows run while the neighbours' boundary values
    arrive by NVLink peer-memory stores (csrc/halo.cu); three-stage SSP-RK3, fixed step;
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


def run(g, steps, alpha=2e-3, a=(0.3, 0.2, 0.1), bw=None, cfl=0.08, combine=True, graph=False):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    dim, n, halo_rows = 3, N_ST, 8
    G = int(round(g * world ** (1.0 / 3.0)))
    shard = rb.SlabShard(rank, world, dim, G, halo_rows)
    ctx = rb.Context(lr, stream=torch.cuda.current_stream().cuda_stream)
    NL, M, o0 = shard.n_local, shard.n_owned, shard.n_lo
    torch.cuda.synchronize()
    t_gen = time.perf_counter()
    X = torch.empty((NL, dim), dtype=torch.float64, device=dev)
    ctx.jittered_lattice_device(dim, G, 0, shard.first_local_id, NL, X.data_ptr())
    own = X[o0:o0 + M]
    st = torch.empty((M, n), dtype=torch.int32, device=dev)
    d2 = torch.empty((M, n), dtype=torch.float64, device=dev)
    ctx.knn_device(X.data_ptr(), NL, dim, n, st.data_ptr(), Q_ptr=own.data_ptr(), NQ=M, d2_out_ptr=d2.data_ptr())
    if not shard.halo_is_sufficient(own[:, -1], d2[:, -1]):
        raise SystemExit("halo too narrow for exact stencils")
    del d2
    opts = rb.make_options(dim, P, n, DEG, OPS)
    parts = []                                      # (row0, row1, operator with the 4 matrices of these rows)
    keep = []
    for (r0, r1) in (rb.boundary_row_ranges(shard) if world > 1 else [(0, M)]):
        if r1 <= r0:
            parts.append((r0, r1, None))
            continue
        ci = torch.empty((r1 - r0, n), dtype=torch.int32, device=dev)
        va = torch.empty((len(OPS), r1 - r0, n), dtype=torch.float64, device=dev)
        ctx.weights_device(opts, X.data_ptr(), NL, st[r0:].data_ptr(), ci.data_ptr(), va.data_ptr(),
                           Y_ptr=own[r0:].data_ptr(), M=r1 - r0, NS=r1 - r0)
        op4 = ctx.operator_from_device(r1 - r0, NL, n, len(OPS), ci.data_ptr(), va.data_ptr())
        if combine:
            # constant coefficients: alpha*Lap - ax*Dx - ay*Dy - az*Dz becomes ONE matrix (the reference rebuilds this sparse
            # sum in every cons_sys call, adv_diff_test.jl:151-152), every stage is then a single-matrix SpMV
            vc = torch.empty((1, r1 - r0, n), dtype=torch.float64, device=dev)
            op4.combine_device([0, 1, 2, 3], [alpha, -a[0], -a[1], -a[2]], vc.data_ptr())
            op4.close()
            del va
            keep.append((ci, vc))
            parts.append((r0, r1, ctx.operator_from_device(r1 - r0, NL, n, 1, ci.data_ptr(), vc.data_ptr())))
        else:
            keep.append((ci, va))
            parts.append((r0, r1, op4))
    del st
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    halo = rb.PeerHalo(ctx, shard) if world > 1 else None
    field = halo.field if halo is not None else torch.zeros(NL, dtype=torch.float64, device=dev)
    field.zero_()
    coef = [1.0] if combine else [alpha, -a[0], -a[1], -a[2]]
    which = [0] if combine else [0, 1, 2, 3]
    du = torch.empty(M, dtype=torch.float64, device=dev)

    def rhs(v):
        """du = alpha Lap v - a . grad v on the owned rows; v = owned values"""
        field[o0:o0 + M].copy_(v)
        if halo is None:
            parts[0][2].spmv_multi_device(which, coef, field.data_ptr(), du.data_ptr())
            return du
        halo.push()
        (l0, l1, opl), (i0, i1, opi), (h0, h1, oph) = parts
        if opi is not None:
            opi.spmv_multi_device(which, coef, field.data_ptr(), du[i0:].data_ptr())     # overlaps the NVLink transfer
        halo.wait()
        if opl is not None:
            opl.spmv_multi_device(which, coef, field.data_ptr(), du[l0:].data_ptr())
        if oph is not None:
            oph.spmv_multi_device(which, coef, field.data_ptr(), du[h0:].data_ptr())
        halo.ack()
        return du

    h = 1.0 / G
    bw = 2.5 * h if bw is None else bw
    inner = ((own > bw) & (own < 1.0 - bw)).all(dim=1)
    bidx = torch.nonzero(~inner).squeeze(1)          # Dirichlet layer
    Xb = own[bidx]
    x0, s0 = (0.35, 0.4, 0.45), max(0.08, 3.0 * h)
    dt = cfl * h * h / alpha
    u = exact(own, 0.0, alpha, a, x0, s0)

    t_dev = torch.zeros((), dtype=torch.float64, device=dev)     # the time lives on the device: the step is replayable

    def stage(v, t):                                 # Dirichlet layer: exact solution at the stage time
        v.index_copy_(0, bidx, exact(Xb, t, alpha, a, x0, s0))
        return v

    def rk3_step():                                  # SSP-RK3 (Shu-Osher)
        v1 = stage(u + dt * rhs(u), t_dev + dt)
        v2 = stage(0.75 * u + 0.25 * (v1 + dt * rhs(v1)), t_dev + 0.5 * dt)
        u.copy_(stage(u / 3.0 + (2.0 / 3.0) * (v2 + dt * rhs(v2)), t_dev + dt))
        t_dev.add_(dt)

    replay, todo = rk3_step, steps
    if graph and world == 1 and steps > 1:
        # one GPU: the step is a fixed sequence of ~30 launches with static buffers -> captured ONCE into a CUDA graph (the
        # context launches on the capture stream) and replayed.  With N > 1 the halo kernels carry an epoch argument that
        # changes every exchange, so the step is not replayable as is.
        rk3_step()
        todo -= 1
        torch.cuda.synchronize()
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
    t0 = time.perf_counter()
    for _ in range(todo):
        replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_run = time.perf_counter() - t0
    t = steps * dt
    steps_timed = max(todo, 1)
    ue = exact(own, t, alpha, a, x0, s0)
    acc = torch.stack([((u - ue) ** 2).sum(), (ue ** 2).sum(), u.sum(), (u * (own[:, 0] + 2 * own[:, 1] + 3 * own[:, 2])).sum()])
    tt = torch.tensor([t_gen, t_run], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(acc)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    err = math.sqrt(float(acc[0]) / float(acc[1]))
    out = {"example": "adv_diff3d_sharded", "n_gpus": world, "global_nodes": G ** 3, "nodes_per_gpu": M, "n": n, "p": P, "polydeg": DEG,
           "steps": steps, "dt": dt, "t_end": t, "rel_l2_error_vs_exact": err, "checksum": [float(acc[2]), float(acc[3])],
           "generation_s": float(tt[0]), "stencils_per_s": G ** 3 / float(tt[0]),
           "ms_per_step": float(tt[1]) / steps_timed * 1e3, "rhs_evaluations_per_s": 3 * steps_timed / float(tt[1]),
           "cuda_graph": bool(graph and world == 1 and steps > 1),
           "operators_per_stage": 1 if combine else 4,
           "halo": "NVLink peer-memory stores (CUDA IPC)" if halo is not None else "none"}
    if halo is not None:
        halo.close()
    return out, u


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--g", type=int, default=48, help="lattice size per GPU: g^3 nodes per rank")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-combine", action="store_true", help="apply the four operators in every stage (fused multi-operator SpMV) "
                    "instead of pre-combining them into one matrix")
    ap.add_argument("--graph", action="store_true", help="one GPU: capture the SSP-RK3 step into a CUDA graph and replay it")
    args = ap.parse_args()
    out, _ = run(args.g, args.steps, combine=not args.no_combine, graph=args.graph)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out))
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
