"""End-to-end host call (rbffd_generate_operator_host, pinned buffers) on configs[1]: ms per call; knobs come from the environment."""
import sys, time
from ctypes import byref, c_void_p
import torch
sys.path.insert(0, ".")
import rbffd_b200 as rb
g, n = 1000, 30
ctx = rb.Context(0)
N = g * g
X = torch.from_numpy(rb.nodes.jittered_lattice(2, g, 0)).pin_memory()
iw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ch = torch.empty((N, n), dtype=torch.int32 if iw == 32 else torch.int64).pin_memory()
vh = torch.empty((1, N, n), dtype=torch.float64).pin_memory()
opts = rb.make_options(2, 5, n, 3, ["Lap"], index_width=iw)
call = lambda: ctx._check(ctx._L.rbffd_generate_operator_host(ctx._h, byref(opts), c_void_p(X.data_ptr()), N, None, N, None, c_void_p(ch.data_ptr()), c_void_p(vh.data_ptr())))
call(); call()
ts = []
for _ in range(7):
    t0 = time.perf_counter(); call(); ts.append(1e3 * (time.perf_counter() - t0))
ts.sort()
print(f"index_width={iw}: median {ts[3]:.2f} ms, min {ts[0]:.2f} ms; weights(device) {ctx.timings()['weights']:.2f} ms")
