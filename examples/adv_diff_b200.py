#!/usr/bin/env python
"""Device-resident advection-diffusion with hyperviscosity: the B200 counterpart of examples/adv_diff_test.jl.

The reference script reads a CGNS mesh, builds E, Dx, Dy, Dxx, Dyy and the hyperviscosity pair with the boundary-aware
methods, and integrates `cons_sys` with SSPRK43.  With `--mesh file.cgns` the node set comes from the same mesh through
rb.mesh.processmesh (BASELINE config 1: examples/rect_0_10.cgns, markers left/right/top/bottom, adv_diff_test.jl:24-29);
without it the node set is a synthetic rectangle [0,5]x[0,1] (interior jittered lattice + boundary midpoints + ghost
nodes offset along the outward normal, the layout src/processmesh.jl:174-186 produces).  Every operator is generated on
the GPU in ONE call (one kNN, one factorisation per node, 7 right-hand sides) and stays in HBM; each RK stage is three
library launches and no torch arithmetic:
    du = cons_sys interior line   (rbffd_rhs_advdiff_device; rows are collocated, so E = I and all six operators are applied
                                   in ONE pass over the shared pattern; --general keeps the E' product of the reference)
    ghost update of the stage     (rbffd_bc_apply_device: all four boundaries in one launch, in the reference's order)
    stage combination             (rbffd_stage_update_device: a*u + b*(v + dt*du), with the UPDATED ghosts as in the reference)
SSP-RK3 with a fixed step stands in for the adaptive SSPRK43 of OrdinaryDiffEq (third-party, not part of the reference package).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbffd_b200 as rb  # noqa: E402


def rectangle_nodes(gy, seed=0):
    """interior jittered lattice on [0,5]x[0,1], boundary midpoints on the 4 sides, ghosts at +0.7h along the normal"""
    gx = 5 * gy
    h = 1.0 / gy
    lin = np.arange(gx * gy)
    u = rb.nodes._splitmix_uniform(seed, lin, 0), rb.nodes._splitmix_uniform(seed, lin, 1)
    Xin = np.stack([((lin % gx) + 0.5 + 0.5 * (u[0] - 0.5)) * h, ((lin // gx) + 0.5 + 0.5 * (u[1] - 0.5)) * h], 1)
    ty, tx = (np.arange(gy) + 0.5) * h, (np.arange(gx) + 0.5) * h
    sides = [(np.stack([np.zeros(gy), ty], 1), (-1, 0)), (np.stack([np.full(gy, 5.0), ty], 1), (1, 0)),
             (np.stack([tx, np.ones(gx)], 1), (0, 1)), (np.stack([tx, np.zeros(gx)], 1), (0, -1))]   # left, right, top, bottom
    bc = [s[0] for s in sides]
    gh = [s[0] + 0.7 * h * np.array(s[1]) for s in sides]
    X = np.concatenate([Xin] + bc + gh)
    n_in = len(Xin)
    sizes = [len(b) for b in bc]
    o = np.concatenate([[n_in], n_in + np.cumsum(sizes)])
    idx_bc = [range(o[b], o[b + 1]) for b in range(4)]
    og = o[-1] + np.concatenate([[0], np.cumsum(sizes)])
    idx_g = [range(og[b], og[b + 1]) for b in range(4)]
    return X, range(0, n_in), idx_bc, idx_g, h


def mesh_nodes(path, ctx=None):
    """adv_diff_test.jl:24-29,78-85: nodes and index sets from the CGNS mesh, h = mean nearest-neighbour distance"""
    X, _, idx_in, idx_bc, idx_g, _, _, _ = rb.mesh.processmesh(path, ["left", "right", "top", "bottom"], ctx=ctx)
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    return X, idx_in, idx_bc, idx_g, float(np.sqrt(d2.min(1)).mean())


def run(gy=40, steps=50, verbose=True, mesh=None, graph=False, collocated=True, timing=None):
    import torch
    dev = torch.device("cuda:0")
    X, idx_in, idx_bc, idx_g, h = mesh_nodes(mesh) if mesh else rectangle_nodes(gy)
    N = len(X)
    p, polydeg = 5, 5
    n = 2 * 21                                                           # adv_diff_test.jl:53-55
    ctx = rb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    names = ["E", "Dx", "Dy", "Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]  # generate_operator + hyperviscosity_operator(2k, ...)
    groups = torch.from_numpy(rb.groups_from_index_sets(N, idx_in, idx_bc, idx_g)).to(dev)
    Xd = torch.from_numpy(X).to(dev)
    t0 = time.perf_counter()
    op = ctx.operator_generate(rb.make_options(2, p, n, polydeg, names), Xd.data_ptr(), N, xgroup_ptr=groups.data_ptr())
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    # ghost updates in the reference's order: right (Dx), left (Dirichlet 1), top (Dy), bottom (Dy)   (:162-176)
    order = [(1, 1), (0, None), (2, 2), (3, 2)]
    bcs = rb.BoundaryConditions(op, [{"bc": list(idx_bc[b]), "ghost": list(idx_g[b]), **({"matrix": m} if m is not None else {"value": 1.0})}
                                     for b, m in order])
    alpha, ux, uy, k = 1.0, 0.0, 0.0, 2                                  # adv_diff_test.jl:88-94
    prm = rb.AdvDiffParams(iE=0, iDx=1, iDy=2, iDxx=3, iDyy=4, iDxk=5, iDyk=6, alpha=alpha, ux=ux, uy=uy, gamma=100.0 * h ** (2 * k),
                           flags=rb.ADVDIFF_COLLOCATED if collocated else 0)
    u = torch.from_numpy(np.where((X[:, 0] - 0.5) ** 2 + (X[:, 1] - 0.5) ** 2 <= 0.04, 10.0, 1.0)).to(dev)   # :65
    du, u1, u2 = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)

    def cons_sys(du_, u_):                                               # adv_diff_test.jl:144-188
        op.rhs_advdiff_device(u_.data_ptr(), du_.data_ptr(), prm)
        bcs.apply_device(u_.data_ptr())

    # explicit stability: the hyperviscosity term 100 h^4 (dx^4 + dy^4) has eigenvalues down to about -700/h^2
    # (the reference leaves the step size to the adaptive SSPRK43 controller)
    dt = 0.0025 * h * h / alpha

    P = lambda t: t.data_ptr()

    def rk3_step():                                                      # SSP-RK3 (Shu-Osher); every line is one library launch
        cons_sys(du, u)
        ctx.stage_update_device(N, 0.0, P(u), 1.0, P(u), dt, P(du), P(u1))              # u1 = u + dt du
        cons_sys(du, u1)
        ctx.stage_update_device(N, 0.75, P(u), 0.25, P(u1), dt, P(du), P(u2))           # u2 = 3/4 u + 1/4 (u1 + dt du)
        cons_sys(du, u2)
        ctx.stage_update_device(N, 1.0 / 3.0, P(u), 2.0 / 3.0, P(u2), dt, P(du), P(u))  # u = 1/3 u + 2/3 (u2 + dt du)

    replay = rk3_step
    if graph and steps > 1:
        # The step is a fixed sequence of 9 small launches over a few thousand nodes (config 1): launch-bound.  One eager
        # step builds the lazily allocated scratch and answers 'is E the identity', then the step is captured ONCE into a CUDA
        # graph on torch's capture stream (the context launches on whatever stream it is given) and replayed.
        rk3_step()
        steps -= 1
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        main_stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.graph(g):
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            rk3_step()
        ctx.set_stream(main_stream)
        replay = g.replay
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        replay()
    torch.cuda.synchronize()
    t_step = (time.perf_counter() - t0) / max(steps, 1)
    uh = u.cpu().numpy()
    if timing is not None:
        timing.update({"ms_per_step": t_step * 1e3, "generation_ms": t_gen * 1e3, "nodes": N, "launches_per_step": 9 if collocated else 15})
    if verbose:
        print(f"N = {N} nodes, n = {n}, generation {t_gen * 1e3:.1f} ms (7 operators), {t_step * 1e3:.3f} ms per SSP-RK3 step "
              f"(3 RHS evaluations{', CUDA graph replay' if graph else ''}), u in [{uh.min():.4f}, {uh.max():.4f}]")
    return X, uh, (idx_in, idx_bc, idx_g)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gy", type=int, default=40)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--mesh", default=None, help="CGNS mesh (e.g. tests/golden/rect_0_10.cgns = BASELINE config 1)")
    ap.add_argument("--graph", action="store_true", help="capture the SSP-RK3 step once into a CUDA graph and replay it")
    ap.add_argument("--general", action="store_true", help="keep the E' product of cons_sys (three products per evaluation) instead of using E = I")
    a = ap.parse_args()
    run(a.gy, a.steps, mesh=a.mesh, graph=a.graph, collocated=not a.general)
