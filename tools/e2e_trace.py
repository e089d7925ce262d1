#!/usr/bin/env python
"""Host-path timeline of rbffd_generate_operator_host at config 2 (RBFFD_TRACE=1 prints the phases to stderr)."""
import os, sys, time
os.environ["RBFFD_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rbffd_b200 as rb
from ctypes import byref, c_void_p
dev = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(dev)
ctx = rb.Context(dev, stream=torch.cuda.current_stream().cuda_stream)
M, n, dim = 1_000_000, 30, 2
X = torch.empty((M, dim), dtype=torch.float64, device="cuda")
ctx.jittered_lattice_device(dim, 1000, 0, 0, M, X.data_ptr())
Xh = torch.empty((M, dim), dtype=torch.float64).pin_memory(); Xh.copy_(X.cpu())
ch = torch.empty((M, n), dtype=torch.int64).pin_memory()
vh = torch.empty((1, M, n), dtype=torch.float64).pin_memory()
opts = rb.make_options(dim, 5, n, 3, ["Lap"])
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx._check(ctx._L.rbffd_generate_operator_host(ctx._h, byref(opts), c_void_p(Xh.data_ptr()), M, None, M, None, c_void_p(ch.data_ptr()), c_void_p(vh.data_ptr())))
    print("[gpu %d] call %d: %.2f ms" % (dev, it, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
