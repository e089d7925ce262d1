#!/usr/bin/env python
"""Forward vs transposed SpMV (D*u vs E'*v, adv_diff_test.jl:151) on a generated operator: time per application and the
fraction of the measured HBM bandwidth by the 12n + 16 B/row convention.  usage: spmv_t_bench.py [dim g]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rbffd_b200 as rb

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
p, n, deg = (5, 30, 3) if dim == 2 else (7, 60, 3)
ctx = rb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
N = g ** dim
X = torch.empty(N, dim, dtype=torch.float64, device="cuda")
ctx.jittered_lattice_device(dim, g, 0, 0, N, X.data_ptr())
op = ctx.operator_generate(rb.make_options(dim, p, n, deg, ["E", "Lap"]), X.data_ptr(), N)
u = torch.randn(N, dtype=torch.float64, device="cuda")
y = torch.empty_like(u)
peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
out = {"dim": dim, "rows": N, "n": n}
for name, fn in (("spmv", lambda: op.spmv_device(1, u.data_ptr(), y.data_ptr())), ("spmv_t", lambda: op.spmv_t_device(1, u.data_ptr(), y.data_ptr()))):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    out[name + "_ms"] = ms
    out[name + "_frac_of_hbm"] = (12 * n + 16) * N / (ms * 1e-3) / 1e9 / peak
print(json.dumps(out))
