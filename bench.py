#!/usr/bin/env python
"""bench.py -- headline benchmark of the RBF-FD hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

A "step" is one pass of the hot path over one batch of synthetic nodes that are already resident in HBM:
    grid-binned exact kNN (stencils + nearest X of every row)  ->  fused weight solve into fixed-row CSR
    ->  one application of the operator (halo exchange of the field first when N > 1).
Workload at N = 1: BASELINE.json configs[1]  "2D Poisson Laplacian operator, 1M synthetic scattered nodes,
PHS r^5 + degree-3 polynomials, k=30".  With N GPUs every rank owns a 1M-node slab of an N-times larger
lattice (weak scaling); weight generation needs no communication, the SpMV needs one halo exchange.

Prints ONE JSON line.  `--impl reference` times the CPU oracle (the stand-in for the reference's Julia CPU path,
which cannot run here: no Julia in the image) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rbffd_stencils_per_s"
UNIT = "stencils/s"
CFG = {"dim": 2, "p": 5, "polydeg": 3, "n": 30, "ops": ["Lap"], "g": 1000}      # configs[1] (the headline workload)
# supplementary workloads (python bench.py --config N): the other BASELINE.json configs at single-GPU sizes
CONFIGS = {
    2: CFG,
    3: {"dim": 2, "p": 5, "polydeg": 4, "n": 50, "ops": ["Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)], "g": 1000},   # configs[2] shape
    4: {"dim": 3, "p": 7, "polydeg": 3, "n": 60, "ops": ["Lap", "Dx", "Dy", "Dz"], "g": 100},                     # configs[3] shape
}


def flops_per_stencil(m, r):
    return (2.0 / 3.0) * m**3 + 2.0 * m * m * r          # getrf + r x getrs convention


def spmv_bytes_per_row(n):
    return 12 * n + 16                                   # fp64 values + int32 indices + x + y


def num_monomials(d, deg):
    c = 1
    for t in range(1, d + 1):
        c = c * (deg + t) // t
    return c


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (profiling recipe, 'clocks line')."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.marks = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """wall-clock marker: samples taken between two marks belong to the timed region"""
        self.marks.append(time.time())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        import datetime
        sm, smax, reasons = [], None, set()
        lo, hi = (self.marks + [0, 0])[0] - 0.03, (self.marks + [0, 0])[1] + 0.03
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                continue
            if len(self.marks) >= 2 and not (lo <= ts <= hi):
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """hardware threads this process may run on (the affinity mask, not OMP_NUM_THREADS)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(sample_g, steps=1, threads=None):
    """Times the CPU oracle (reference arithmetic: kd-tree kNN, inv(A)*RHS per stencil, CSR SpMV) on a g^2 sample."""
    import numpy as np
    import rbffd_b200 as rb
    from oracle import oracle as orc
    # all host cores, whatever the launcher exported: torch.distributed.run sets OMP_NUM_THREADS=1 for its workers
    orc.set_num_threads(threads or host_cores())
    X = rb.nodes.jittered_lattice(CFG["dim"], sample_g, 0)
    N = len(X)
    best = None
    u = np.random.default_rng(0).standard_normal(N)
    for _ in range(steps):
        t0 = time.perf_counter()
        colind, vals = orc.generate_operator(X, X, CFG["p"], CFG["n"], CFG["polydeg"], ops=CFG["ops"], mode=0)
        orc.spmv(colind, vals[0], u)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return N / best, best, N, orc.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_g = args.ref_sample_g
    times = []
    value = 0.0
    for i in range(args.warmup + args.steps):
        v, dt, Ns, cores = cpu_sample(sample_g)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = Ns / (ms * 1e-3)
    m = CFG["n"] + num_monomials(CFG["dim"], CFG["polydeg"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 2D Laplacian operator, jittered lattice, PHS r^5 + deg-3 polynomials, k=30 (m=%d)" % m,
                   "sample": f"{Ns} of 1000000 nodes per step", "algorithm": "CPU restatement of the reference: kd-tree kNN, "
                   "LU+inverse (getrf/getri) per stencil, inv(A)*RHS, CSR SpMV; C + OpenMP (oracle/rbffd_oracle.c)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{Ns}-node sample (g={sample_g}) of the 1M-node workload, all host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is pure Julia and Julia is not installed on the box: the C/OpenMP oracle stands in for it",
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--g", type=int, default=CFG["g"], help="lattice size per GPU: g^2 nodes per rank")
    ap.add_argument("--ref-sample-g", type=int, default=400)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4], help="BASELINE.json config shape (2 = headline)")
    ap.add_argument("--profile", action="store_true", help="ncu passes: honour small --warmup, skip e2e and the CPU baseline")
    args = ap.parse_args()
    if args.config != 2:
        CFG.update(CONFIGS[args.config])
        if args.g == 1000:
            args.g = CFG["g"]
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import rbffd_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    W = args.warmup if args.profile else max(args.warmup, 3)
    K = args.steps

    dim, p, deg, n = CFG["dim"], CFG["p"], CFG["polydeg"], CFG["n"]
    q = num_monomials(dim, deg)
    m = n + q
    r = len(CFG["ops"])
    # global lattice: world * g^2 nodes, G x G with G = round(g * sqrt(world)); rank owns a slab of rows
    G = int(round(args.g * world ** (1.0 / dim)))
    halo_rows = 8
    shard = rb.SlabShard(rank, world, dim, G, halo_rows)
    stream = torch.cuda.current_stream()
    ctx = rb.Context(local_rank, stream=stream.cuda_stream)
    peak = ctx.measure_fp64_peak()

    NL, M = shard.n_local, shard.n_owned
    X = torch.empty((NL, dim), dtype=torch.float64, device=dev)
    ctx.jittered_lattice_device(dim, G, 0, shard.first_local_id, NL, X.data_ptr())
    Yown = X[shard.n_lo:shard.n_lo + M]                       # owned nodes = rows of the operator (Y == X, collocated)
    opts = rb.make_options(dim, p, n, deg, CFG["ops"], kernel=args.kernel)
    stencils = torch.empty((M, n), dtype=torch.int32, device=dev)
    d2 = torch.empty((M, n), dtype=torch.float64, device=dev) if world > 1 else None
    center = torch.empty(M, dtype=torch.int32, device=dev)
    colind = torch.empty((M, n), dtype=torch.int32, device=dev)
    vals = torch.empty((r, M, n), dtype=torch.float64, device=dev)
    # N > 1: the field lives in a CUDA-IPC buffer; neighbours store their boundary rows straight into its halos over
    # NVLink (csrc/halo.cu).  RBFFD_HALO=nccl selects the torch.distributed send/recv path instead.
    use_p2p = world > 1 and os.environ.get("RBFFD_HALO", "p2p") != "nccl"
    halo = rb.PeerHalo(ctx, shard) if use_p2p else None
    u = halo.field if use_p2p else torch.empty(NL, dtype=torch.float64, device=dev)
    u.copy_(torch.randn(NL, dtype=torch.float64, device=dev))
    y = torch.empty(M, dtype=torch.float64, device=dev)
    op = ctx.operator_from_device(M, NL, n, r, colind.data_ptr(), vals.data_ptr())
    # N > 1: rows that cannot reference halo columns are applied while the halo exchange is in flight
    parts = []
    if world > 1:
        for (r0, r1) in rb.boundary_row_ranges(shard):
            parts.append((r0, r1, ctx.operator_from_device(r1 - r0, NL, n, 1, colind[r0:].data_ptr(), vals[0, r0:].data_ptr()) if r1 > r0 else None))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    phase = {"knn": 0.0, "weights": 0.0, "spmv": 0.0}

    def step(record):
        ev[0].record(stream)
        if world == 1:
            # stencils of every node + nearest X of every row, one binning pass (generate_operator.jl:43-47)
            ctx.stencils_device(X.data_ptr(), NL, dim, n, stencils.data_ptr(), center_ptr=center.data_ptr())
        else:
            ctx.knn_device(X.data_ptr(), NL, dim, n, stencils.data_ptr(), Q_ptr=Yown.data_ptr(), NQ=M, d2_out_ptr=d2.data_ptr())
        ev[1].record(stream)
        ctx.weights_device(opts, X.data_ptr(), NL, stencils.data_ptr(), colind.data_ptr(), vals.data_ptr(),
                           Y_ptr=Yown.data_ptr(), M=M, center_ptr=center.data_ptr() if world == 1 else None, NS=M)
        ev[2].record(stream)
        if world > 1:
            work = None
            if use_p2p:
                halo.push()
            else:
                work = rb.exchange_halo(u, shard, async_op=True)
            (l0, l1, opl), (i0, i1, opi), (h0, h1, oph) = parts
            if opi is not None:
                opi.spmv_device(0, u.data_ptr(), y[i0:].data_ptr())          # interior rows overlap the NVLink transfer
            if use_p2p:
                halo.wait()
            else:
                work.wait()
            if opl is not None:
                opl.spmv_device(0, u.data_ptr(), y[l0:].data_ptr())
            if oph is not None:
                oph.spmv_device(0, u.data_ptr(), y[h0:].data_ptr())
            if use_p2p:
                halo.ack()
        else:
            op.spmv_device(0, u.data_ptr(), y.data_ptr())
        ev[3].record(stream)
        if record:
            ev[3].synchronize()
            phase["knn"] += ev[0].elapsed_time(ev[1])
            phase["weights"] += ev[1].elapsed_time(ev[2])
            phase["spmv"] += ev[2].elapsed_time(ev[3])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()           # started before the warm-up so nvidia-smi is already sampling when the timed region begins
    for _ in range(W):
        step(False)
    barrier()
    if world > 1:
        ok = shard.halo_is_sufficient(Yown[:, -1], d2[:, -1])
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() != 1:
            raise SystemExit("halo too narrow for exact stencils: increase halo_rows")
        # rows applied during the exchange must not reference halo columns (exactness of the overlap)
        (i0, i1) = rb.boundary_row_ranges(shard)[1]
        if i1 > i0:
            ci = colind[i0:i1]
            if int(ci.min()) < shard.n_lo or int(ci.max()) >= shard.n_lo + shard.n_owned:
                raise SystemExit("interior rows reference halo columns: widen the boundary row ranges")
    launches0 = ctx.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark()
    t_start.record(stream)
    for _ in range(K):
        step(True)
    t_end.record(stream)
    barrier()
    clocks.mark()
    elapsed_ms = t_start.elapsed_time(t_end)
    launches = ctx.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / K
    total_nodes = G ** dim
    value = total_nodes / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing host call: pinned host X in, host CSR out, copies inside the timing
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_per_step, "phases_ms": {k: v / K for k, v in phase.items()}}))
        if halo is not None:
            halo.close()
        if world > 1:
            dist.destroy_process_group()
        return
    Ke = max(1, min(args.e2e_steps, K))
    Xh = torch.empty((M, dim), dtype=torch.float64).pin_memory()
    Xh.copy_(Yown.cpu())
    ch = torch.empty((M, n), dtype=torch.int64).pin_memory()
    vh = torch.empty((r, M, n), dtype=torch.float64).pin_memory()
    from ctypes import byref, c_void_p
    L = ctx._L
    def e2e_step():
        ctx._check(L.rbffd_generate_operator_host(ctx._h, byref(opts), c_void_p(Xh.data_ptr()), M, None, M, None,
                                                  c_void_p(ch.data_ptr()), c_void_p(vh.data_ptr())))
    e2e_step()
    e2e_step()
    barrier()
    # the call blocks until the caller's host buffers hold the result, so wall clock around it is the end-to-end time;
    # no collective inside the timed region (ranks are independent here), the max over ranks is taken below
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / Ke
    barrier()
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = (M * world) / e2e_s
    h2d = M * dim * 8
    d2h = M * n * 8 + r * M * n * 8

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        t_w = phase["weights"] / K * 1e-3
        t_s = phase["spmv"] / K * 1e-3
        t_k = phase["knn"] / K * 1e-3
        F = flops_per_stencil(m, r)
        fp64_peak = max(peak["dfma_tflops"], peak["dmma_tflops"])
        ach_w = F * M / t_w * 1e-12
        ach_s = spmv_bytes_per_row(n) * M / t_s * 1e-9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("configs[1]: 2D Laplacian operator, %d jittered-lattice nodes per GPU, PHS r^5 + deg-3 polynomials, "
                                    "k=30 (m=%d, r=%d)" % (M, m, r)) if args.config == 2 else
                                   ("configs[%d] shape: %dD, %d nodes per GPU, p=%d, polydeg=%d, k=%d (m=%d, r=%d) ops=%s"
                                    % (args.config - 1, dim, M, p, deg, n, m, r, CFG["ops"])),
                       "step": "exact kNN + fused weight solve -> CSR + one SpMV (halo exchange first if N>1); nodes resident in HBM",
                       "l2": "every step writes %.0f MB of stencils+operator (> 126 MB L2), so no input survives in L2 between steps" % ((M * n * 4 * 2 + r * M * n * 8) / 1e6),
                       "parallelism": "slab x%d, halo_rows=%d, halo exchange: %s" % (world, halo_rows, "none" if world == 1 else ("NVLink peer-memory stores (CUDA IPC)" if use_p2p else "NCCL send/recv")), "global_nodes": total_nodes},
            "phases_ms": {"knn": phase["knn"] / K, "weights": phase["weights"] / K, "spmv(+halo)": phase["spmv"] / K},
            "roofline": {"kernel": "fused weight kernel (assemble + null-space elimination + solve + CSR write), flops by the LU convention (2/3)m^3 + 2m^2 r", "bound": "fp64",
                         "achieved": ach_w, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_w / fp64_peak,
                         "flop_per_stencil": F, "traffic": traffic.get("weights"),
                         "peak_source": "measured live in this run by rbffd_measure_fp64_peak (DFMA %.2f / DMMA %.2f TFLOP/s burst); "
                                        "MEASURED_PEAKS.json has no FP64 entry" % (peak["dfma_tflops"], peak["dmma_tflops"])},
            "roofline_spmv": {"kernel": "spmv_multi_kernel", "bound": "hbm", "achieved": ach_s, "peak": hbm_peak, "unit": "GB/s",
                              "frac": ach_s / hbm_peak, "bytes_per_row": spmv_bytes_per_row(n), "traffic": traffic.get("spmv"),
                              "peak_source": hbm_src, "note": "includes the halo exchange when N>1"},
            "knn": {"queries_per_s": M / t_k, "ms": t_k * 1e3},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": Ke,
                    "pcie_d2h_bytes_per_step": M * n * 4 + r * M * n * 8,
                    "call": "rbffd_generate_operator_host (pinned host X in, int64 colind + fp64 values out)"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, dt, Ns, cores = cpu_sample(args.ref_sample_g)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{Ns}-node sample (g={args.ref_sample_g}) of the same workload, {dt:.2f} s; CPU oracle = "
                                              "C/OpenMP restatement of the reference (no Julia on the box)"}
        print(json.dumps(line))
    if halo is not None:
        halo.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
