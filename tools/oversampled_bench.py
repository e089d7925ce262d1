#!/usr/bin/env python
"""Oversampled mode (Y != X, generate_operator.jl:89-167; poisson_test.jl:56 uses M ~ 3N): weight-phase time of the generic
kernel (one factorisation per centre, kernel=1) against the row-wise null-space kernels (default dispatch), device resident.
usage: oversampled_bench.py [dim g over [lap]]   (lap: one operator, the Laplacian, instead of the reference tuple)"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import rbffd_b200 as rb

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
over = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p, n, deg, ops = (5, 30, 3, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"]) if dim == 2 else (7, 60, 3, ["Lap", "Dx", "Dy", "Dz"])
if len(sys.argv) > 4 and sys.argv[4] == "lap":
    ops = ["Lap"]
ctx = rb.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
N = g ** dim
X = torch.empty(N, dim, dtype=torch.float64, device="cuda")
ctx.jittered_lattice_device(dim, g, 0, 0, N, X.data_ptr())
M = over * N
Y = torch.rand(M, dim, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1)) * 0.96 + 0.02
out = {"dim": dim, "N": N, "M": M, "p": p, "n": n, "polydeg": deg, "ops": [str(o) for o in ops]}
for name, kernel in (("generic_kernel1", 1), ("rowwise_nullspace", 0)):
    opts = rb.make_options(dim, p, n, deg, ops, 0, False, kernel, 0)
    best = None
    for it in range(3):
        op = ctx.operator_generate(opts, X.data_ptr(), N, Y.data_ptr(), M)
        t = ctx.timings()
        best = t["weights"] if best is None else min(best, t["weights"])
        if it == 2:
            ci, vi = op.pointers(0)
        op.close()
    out[name + "_weights_ms"] = best
    out[name + "_rows_per_s"] = M / best * 1e3
out["speedup"] = out["generic_kernel1_weights_ms"] / out["rowwise_nullspace_weights_ms"]
print(json.dumps(out))
