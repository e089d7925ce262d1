"""kNN at BASELINE configs[3] full size (3-D, 19.9 M nodes, k = 60): one stencils_device call (profiling helper)."""
import sys, time
import torch
sys.path.insert(0, ".")
import rbffd_b200 as rb
g = int(sys.argv[1]) if len(sys.argv) > 1 else 271
k = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dim = 3
dev = torch.device("cuda:0")
ctx = rb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
N = g ** dim
X = torch.empty((N, dim), dtype=torch.float64, device=dev)
ctx.jittered_lattice_device(dim, g, 0, 0, N, X.data_ptr())
st = torch.empty((N, k), dtype=torch.int32, device=dev)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.stencils_device(X.data_ptr(), N, dim, k, st.data_ptr())
    torch.cuda.synchronize(); print(f"g={g} N={N} k={k}: {1e3 * (time.perf_counter() - t0):.1f} ms, timings {ctx.timings()}")
