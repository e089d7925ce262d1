"""Multi-GPU check of the NVLink peer-memory halo exchange (run under torchrun on >= 2 GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_halo_check.py
Verifies (1) halo contents after push/wait/ack over many epochs, (2) the sharded SpMV with overlapped peer-memory halos
against the NCCL-exchange path and against a single-GPU application of the global operator rows, and prints timings."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbffd_b200 as rb  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    g = int(os.environ.get("HALO_G", "600"))
    dim, n, halo_rows = 2, 30, 8
    G = int(round(g * world ** 0.5))
    s = rb.SlabShard(rank, world, dim, G, halo_rows)
    stream = torch.cuda.current_stream()
    ctx = rb.Context(lr, stream=stream.cuda_stream)
    halo = rb.PeerHalo(ctx, s)
    u = halo.field
    gid = torch.arange(s.first_local_id, s.first_local_id + s.n_local, dtype=torch.float64, device=dev)
    o0, o1 = s.n_lo, s.n_lo + s.n_owned
    ok = True
    for ep in range(1, 41):
        u.fill_(-1.0)                      # test only: the owner never writes its halos in real use
        u[o0:o1] = gid[o0:o1] * ep + 0.5
        torch.cuda.synchronize()
        dist.barrier()                     # ... so keep the neighbours' pushes behind this rank's own fill
        halo.push()
        halo.wait()
        good = torch.equal(u, gid * ep + 0.5)
        halo.ack()
        torch.cuda.synchronize()
        dist.barrier()          # the test rewrites the halos itself (fill_), so keep the epochs apart
        ok = ok and bool(good)
    ok_epochs = ok
    # sharded operator on this slab
    X = torch.empty((s.n_local, dim), dtype=torch.float64, device=dev)
    ctx.jittered_lattice_device(dim, G, 0, s.first_local_id, s.n_local, X.data_ptr())
    own = X[o0:o1]
    M = s.n_owned
    st = torch.empty((M, n), dtype=torch.int32, device=dev)
    colind = torch.empty((M, n), dtype=torch.int32, device=dev)
    vals = torch.empty((1, M, n), dtype=torch.float64, device=dev)
    ctx.knn_device(X.data_ptr(), s.n_local, dim, n, st.data_ptr(), Q_ptr=own.data_ptr(), NQ=M)
    opts = rb.make_options(dim, 5, n, 3, ["Lap"])
    ctx.weights_device(opts, X.data_ptr(), s.n_local, st.data_ptr(), colind.data_ptr(), vals.data_ptr(), Y_ptr=own.data_ptr(), M=M, NS=M)
    parts = [(r0, r1, ctx.operator_from_device(r1 - r0, s.n_local, n, 1, colind[r0:].data_ptr(), vals[0, r0:].data_ptr()) if r1 > r0 else None)
             for (r0, r1) in rb.boundary_row_ranges(s)]
    full = ctx.operator_from_device(M, s.n_local, n, 1, colind.data_ptr(), vals.data_ptr())
    f = torch.sin(3 * X[:, 0]) * torch.cos(2 * X[:, 1])          # field values at every local node (owned + halo): the truth
    y_ref = torch.empty(M, dtype=torch.float64, device=dev)
    full.spmv_device(0, f.data_ptr(), y_ref.data_ptr())

    def apply_p2p(y):
        halo.push()
        (l0, l1, opl), (i0, i1, opi), (h0, h1, oph) = parts
        if opi is not None:
            opi.spmv_device(0, u.data_ptr(), y[i0:].data_ptr())
        halo.wait()
        if opl is not None:
            opl.spmv_device(0, u.data_ptr(), y[l0:].data_ptr())
        if oph is not None:
            oph.spmv_device(0, u.data_ptr(), y[h0:].data_ptr())
        halo.ack()

    def apply_nccl(y, uu):
        rb.exchange_halo(uu, s)
        full.spmv_device(0, uu.data_ptr(), y.data_ptr())

    u.fill_(0.0)
    u[o0:o1] = f[o0:o1]
    torch.cuda.synchronize()
    dist.barrier()
    y1 = torch.empty(M, dtype=torch.float64, device=dev)
    apply_p2p(y1)
    u2 = torch.zeros_like(f)
    u2[o0:o1] = f[o0:o1]
    y2 = torch.empty(M, dtype=torch.float64, device=dev)
    apply_nccl(y2, u2)
    torch.cuda.synchronize()
    e1 = float((y1 - y_ref).abs().max()); e2 = float((y2 - y_ref).abs().max())
    print(f'[rank {rank}] epochs ok={ok_epochs} p2p err={e1:.3e} nccl err={e2:.3e} ref max={float(y_ref.abs().max()):.3e}', flush=True)
    ok = ok and torch.equal(y1, y_ref) and torch.equal(y2, y_ref)
    # timings
    res = {}
    for name, fn in (("p2p", lambda: apply_p2p(y1)), ("nccl", lambda: apply_nccl(y2, u2))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(50):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        res[name] = (time.perf_counter() - t0) / 50 * 1e3
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"mgpu_halo_check world={world} rows/rank={M}: {'OK' if flag.item() == 1 else 'FAILED'}; "
              f"SpMV + halo per application: peer-memory {res['p2p']:.3f} ms, NCCL send/recv {res['nccl']:.3f} ms")
    halo.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 3)


if __name__ == "__main__":
    main()
