#!/usr/bin/env python
"""bench.py -- headline benchmark of the RBF-FD hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

A "step" is one pass of the hot path over one batch of synthetic nodes that are already resident in HBM:
    grid-binned exact kNN (stencils + nearest X of every row)  ->  fused weight solve into fixed-row CSR
    ->  one application of the operator (its halo exchange fused into the same launch when N > 1).
Workload at N = 1: BASELINE.json configs[1]  "2D Poisson Laplacian operator, 1M synthetic scattered nodes,
PHS r^5 + degree-3 polynomials, k=30".  With N GPUs every rank owns a 1M-node spatial block of an N-times larger
lattice (weak scaling; blocks 2x1 / 2x2 / 4x2); weight generation needs no communication, the SpMV needs one halo
exchange, and the sharded product is CHECKED against the truth in every run (the run fails otherwise).
The same JSON line carries a `configs` block: the full-size BASELINE configs[2] (10M nodes) and configs[3] (20M nodes)
at N = 1, configs[4] (12.5M nodes per GPU: 100M on 8 GPUs, sharded generation + SSP-RK3 time stepping) at N > 1, and the
end-to-end host call for the reference's six-operator tuple.

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


def pseudo_field(gid):
    """closed-form field value of a node from its GLOBAL id: every rank can evaluate the true halo values without communication"""
    return ((gid * 2654435761) % 1000003).double() / 1000003.0 - 0.5


def full_size_config(rb, ctx, torch, dev, cfg, g, peak, hbm_peak, passes=3):
    """one BASELINE config at its full single-GPU size: kNN, weights, SpMV (ms per pass), both roofline fractions.

    One untimed pass, then `passes` timed ones; the reported time of a phase is the MEDIAN over the passes and the
    individual passes are listed in `per_pass_ms`: on some boxes single cudaMallocAsync calls were seen to block for
    50-400 ms (profiles/r02ah_fullsize_pass_variance.txt, r02aq_alloc_stalls.txt); the library parks its temporaries in
    a host-side free list since then and the passes are steady -- the median and the list stay as the guard."""
    dim, p, deg, n, ops = cfg["dim"], cfg["p"], cfg["polydeg"], cfg["n"], cfg["ops"]
    N = g ** dim
    r = len(ops)
    m = n + num_monomials(dim, deg)
    stream = torch.cuda.current_stream()
    X = torch.empty((N, dim), dtype=torch.float64, device=dev)
    ctx.jittered_lattice_device(dim, g, 0, 0, N, X.data_ptr())
    opts = rb.make_options(dim, p, n, deg, ops)
    stencils = torch.empty((N, n), dtype=torch.int32, device=dev)
    center = torch.empty(N, dtype=torch.int32, device=dev)
    colind = torch.empty((N, n), dtype=torch.int32, device=dev)
    vals = torch.empty((r, N, n), dtype=torch.float64, device=dev)
    u = torch.randn(N, dtype=torch.float64, device=dev)
    y = torch.empty(N, dtype=torch.float64, device=dev)
    op = ctx.operator_from_device(N, N, n, r, colind.data_ptr(), vals.data_ptr())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = [0.0, 0.0, 0.0]
    per_pass = {"knn": [], "weights": [], "spmv": []}
    inner = {"binning": [], "knn": [], "nearest": []}
    for it in range(passes + 1):
        ev[0].record(stream)
        ctx.stencils_device(X.data_ptr(), N, dim, n, stencils.data_ptr(), center_ptr=center.data_ptr())
        ev[1].record(stream)
        if it > 0:
            tm = ctx.timings()
            for kk in inner:
                inner[kk].append(tm[kk])
        ctx.weights_device(opts, X.data_ptr(), N, stencils.data_ptr(), colind.data_ptr(), vals.data_ptr(), Y_ptr=X.data_ptr(), M=N,
                           center_ptr=center.data_ptr(), NS=N)
        ev[2].record(stream)
        op.spmv_device(0, u.data_ptr(), y.data_ptr())
        ev[3].record(stream)
        ev[3].synchronize()
        if it > 0:
            for k, name in enumerate(("knn", "weights", "spmv")):
                per_pass[name].append(round(ev[k].elapsed_time(ev[k + 1]), 3))
    median = lambda v: sorted(v)[len(v) // 2]
    t = [median(per_pass[name]) for name in ("knn", "weights", "spmv")]
    inner = {kk: median(v) for kk, v in inner.items()}
    F = flops_per_stencil(m, r)
    fp64_peak = max(peak["dfma_tflops"], peak["dmma_tflops"])
    ach_w = F * N / (t[1] * 1e-3) * 1e-12
    ach_s = spmv_bytes_per_row(n) * N / (t[2] * 1e-3) * 1e-9
    # sanity of the result at full size: rows of a derivative operator sum to zero, weights finite
    rs = vals[0].sum(dim=1).abs().max().item() / vals[0].abs().sum(dim=1).max().item()
    out = {"nodes": N, "dim": dim, "p": p, "polydeg": deg, "n": n, "m": m, "r": r, "ops": [str(o) for o in ops], "passes": passes,
           "knn_ms": t[0], "knn_breakdown_ms": inner, "weights_ms": t[1], "spmv_ms": t[2], "stencils_per_s": N / ((t[0] + t[1] + t[2]) * 1e-3),
           "roofline_weights": {"achieved": ach_w, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_w / fp64_peak, "flop_per_stencil": F},
           "roofline_spmv": {"achieved": ach_s, "peak": hbm_peak, "unit": "GB/s", "frac": ach_s / hbm_peak, "bytes_per_row": spmv_bytes_per_row(n)},
           "row_sum_defect": rs, "per_pass_ms": per_pass}
    op.close()
    del X, stencils, center, colind, vals, u, y
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (full-size configs 3/4 at N=1, config 5 at N>1)")
    ap.add_argument("--cfg5-g", type=int, default=232, help="configs[4]: lattice size per GPU (232^3 = 12.5M nodes; 100M on 8 GPUs)")
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
    # stdout carries exactly ONE line (the JSON): anything a library prints there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import numpy as np
    import torch
    import torch.distributed as dist
    import rbffd_b200 as rb
    from rbffd_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    # one process per GPU: stay on the cores next to this GPU (pinned staging buffers and the host threads of the end-to-end
    # call then live on its NUMA node)
    bound_cpus = rb.bind_to_gpu_numa(local_rank) if world > 1 else 0
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
    # global lattice: world * g^dim nodes, cut into one spatial block per rank
    G = int(round(args.g * world ** (1.0 / dim)))
    blocks = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (4, 2, 1) if dim == 2 else (2, 2, 2)}.get(world)
    if blocks is None:
        raise SystemExit("bench.py runs on 1, 2, 4 or 8 GPUs")
    blocks = blocks[:dim]
    stream = torch.cuda.current_stream()
    ctx = rb.Context(local_rank, stream=stream.cuda_stream)
    peak = ctx.measure_fp64_peak()
    opts = rb.make_options(dim, p, n, deg, CFG["ops"], kernel=args.kernel)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    phase = {"knn": 0.0, "weights": 0.0, "spmv": 0.0}
    shard = None
    t_shard = 0.0
    # RBFFD_HALO=nccl: halo values travel by torch.distributed send/recv instead of peer-memory stores inside the product launch
    use_nccl = world > 1 and os.environ.get("RBFFD_HALO", "p2p") == "nccl"

    if world == 1:
        NL = M = G ** dim
        X = torch.empty((NL, dim), dtype=torch.float64, device=dev)
        ctx.jittered_lattice_device(dim, G, 0, 0, NL, X.data_ptr())
        X_ptr, Q_ptr = X.data_ptr(), X.data_ptr()
        center = torch.empty(M, dtype=torch.int32, device=dev)
    else:
        # spatial-block shard: owned nodes + halo (stencil closure), local numbering [interior | boundary | halo]; built once
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        shard = sharding.Shard.lattice_block(ctx, dim, G, 0, blocks, rank, n)
        shard.wire(ipc=not use_nccl)
        torch.cuda.synchronize()
        t_shard = time.perf_counter() - t0
        M, NL = shard.n_owned, shard.n_owned + shard.n_halo
        X_ptr = Q_ptr = shard.X_local_ptr
        center = None
    stencils = torch.empty((M, n), dtype=torch.int32, device=dev)
    colind = torch.empty((M, n), dtype=torch.int32, device=dev)
    vals = torch.empty((r, M, n), dtype=torch.float64, device=dev)
    u = torch.randn(M, dtype=torch.float64, device=dev)
    y = torch.empty(M, dtype=torch.float64, device=dev)
    op = ctx.operator_from_device(M, NL, n, r, colind.data_ptr(), vals.data_ptr())

    def sharded_apply(o, xin, yout):
        if use_nccl:
            shard.exchange_collective(xin)                                     # RBFFD_HALO=nccl: pack -> NCCL send/recv -> unpack
            shard.spmv_local_device(o, [0], [1.0], xin.data_ptr(), yout.data_ptr())
        else:
            shard.spmv_device(o, [0], [1.0], xin.data_ptr(), yout.data_ptr())  # ONE launch: push + interior rows + wait + boundary rows + ack

    def step(record):
        ev[0].record(stream)
        if world == 1:
            # stencils of every node + nearest X of every row, one binning pass (generate_operator.jl:43-47)
            ctx.stencils_device(X_ptr, NL, dim, n, stencils.data_ptr(), center_ptr=center.data_ptr())
        else:
            # the rank's own search: owned nodes among [owned | halo]; no communication
            ctx.knn_device(X_ptr, NL, dim, n, stencils.data_ptr(), Q_ptr=Q_ptr, NQ=M)
        ev[1].record(stream)
        ctx.weights_device(opts, X_ptr, NL, stencils.data_ptr(), colind.data_ptr(), vals.data_ptr(),
                           Y_ptr=Q_ptr, M=M, center_ptr=center.data_ptr() if world == 1 else None, NS=M)
        ev[2].record(stream)
        if world > 1:
            sharded_apply(op, u, y)
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
    parity = None
    if world > 1:
        # ---- parity of the sharded path, checked in EVERY run: (1) the stencils found in the timed step are the shard's
        # (global-id tie-broken) stencils, (2) the fused halo-exchange SpMV equals the plain product of the same rows over
        # [owned values ; TRUE halo values], the truth coming from a closed-form field of the global ids (no communication)
        gid = torch.from_numpy(shard.global_ids()).to(dev)
        ref_st = torch.empty((M, n), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        # device-to-device copy of the shard's stencils through a zero-copy view
        class _Raw:
            __cuda_array_interface__ = {"shape": (M, n), "typestr": "<i4", "data": (shard.stencils_ptr, False), "version": 3, "strides": None}
        ref_st.copy_(torch.as_tensor(_Raw(), device=dev))
        same_sets = bool((torch.sort(stencils, dim=1).values == torch.sort(ref_st, dim=1).values).all())
        full = pseudo_field(gid)
        u.copy_(full[:M])
        y_ref = torch.empty(M, dtype=torch.float64, device=dev)
        for rep in range(3):                                   # several epochs back to back
            sharded_apply(op, u, y)
        op.spmv_multi_device([0], [1.0], full.data_ptr(), y_ref.data_ptr())
        torch.cuda.synchronize()
        identical = bool(torch.equal(y, y_ref))
        maxdiff = float((y - y_ref).abs().max())
        flag = torch.tensor([1 if (identical and same_sets) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"stencils_equal_global_id_stencils": same_sets, "sharded_spmv_bit_identical_to_truth": identical, "max_abs_diff": maxdiff,
                  "all_ranks_ok": bool(flag.item() == 1)}
        if flag.item() != 1:
            raise SystemExit(f"rank {rank}: sharded SpMV / stencils differ from the truth: {parity}")
        del ref_st, gid, full, y_ref
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
    if shard is not None:
        shard.check()                # a halo exchange that timed out (dead peer) is an error, not a number
    elapsed_ms = t_start.elapsed_time(t_end)
    launches = ctx.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    total_nodes = M
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        c = torch.tensor([float(M)], device=dev, dtype=torch.float64)
        dist.all_reduce(c)
        total_nodes = int(c.item())
    ms_per_step = elapsed_ms / K
    value = total_nodes / (ms_per_step * 1e-3)
    # ---- the operator application on its own: a loop of back-to-back products (the time-stepping regime).  Inside the step the
    # product follows a 5 ms weight solve whose duration differs from rank to rank, so its in-step time at N > 1 contains that
    # skew (the boundary rows wait for the slowest neighbour's values); the loop measures the product with its halo exchange.
    # A ring of three copies of the operator is cycled through, so every application streams its matrix from HBM (one copy
    # is 360 MB; a single copy applied in a loop would keep a third of itself in the 126 MB L2).
    y2 = torch.empty_like(y)
    reps = 48
    ring = [(op, None)]
    for _ in range(2):
        ci2, va2 = colind.clone(), vals[0].clone()
        ring.append((ctx.operator_from_device(M, NL, n, 1, ci2.data_ptr(), va2.data_ptr()), (ci2, va2)))

    def apply(o):
        if world > 1:
            sharded_apply(o, u, y2)
        else:
            o.spmv_device(0, u.data_ptr(), y2.data_ptr())
    for i in range(6):
        apply(ring[i % 3][0])
    barrier()
    sp0, sp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sp0.record(stream)
    for i in range(reps):
        apply(ring[i % 3][0])
    sp1.record(stream)
    barrier()
    spmv_loop_ms = sp0.elapsed_time(sp1) / reps
    if world > 1:
        t = torch.tensor([spmv_loop_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        spmv_loop_ms = float(t.item())
    for o, keep in ring[1:]:
        o.close()
    del y2, ring

    if args.profile:
        if rank == 0:
            emit({"profile_run": True, "ms_per_step": ms_per_step, "phases_ms": {k: v / K for k, v in phase.items()}})
        if world > 1:
            shard.close()
            dist.destroy_process_group()
        return

    # ---- end to end through the reference-facing host call: pinned host buffers in, host CSR out, copies inside the timing
    Ke = max(1, min(args.e2e_steps, K))
    from ctypes import byref, c_void_p
    L = ctx._L
    if world == 1:
        Xh = torch.empty((M, dim), dtype=torch.float64).pin_memory()
        Xh.copy_(X.cpu())
        ch = torch.empty((M, n), dtype=torch.int64).pin_memory()
        vh = torch.empty((r, M, n), dtype=torch.float64).pin_memory()

        def e2e_step():
            ctx._check(L.rbffd_generate_operator_host(ctx._h, byref(opts), c_void_p(Xh.data_ptr()), M, None, M, None,
                                                      c_void_p(ch.data_ptr()), c_void_p(vh.data_ptr())))
        h2d = M * dim * 8
        d2h = M * n * 8 + r * M * n * 8
        e2e_call = "rbffd_generate_operator_host (pinned host X in, int64 colind + fp64 values out)"
    else:
        # the rank's part of the sharded operator through the SAME drop-in host call: the rank hands over its block plus the
        # candidate margin (coordinates in ascending global id) and keeps the rows of its owned nodes; column ids refer to the
        # candidate numbering (the caller holds the global ids).  No new API, no communication: this is how a Julia caller
        # shards generate_operator over the GPUs of one box.  The rows it keeps are CHECKED against the shard's rows below.
        import math
        import ctypes as C
        b, rr = [0] * dim, rank
        for a in reversed(range(dim)):
            b[a] = rr % blocks[a]
            rr //= blocks[a]
        mg = shard.margin
        lo = [max(0, -((-b[a] * G) // blocks[a]) - mg) for a in range(dim)]
        hi = [min(G, -((-(b[a] + 1) * G) // blocks[a]) + mg) for a in range(dim)]
        nc = math.prod(h - l for l, h in zip(lo, hi))
        Xc = torch.empty(nc * dim, dtype=torch.float64, device=dev)
        gc = torch.empty(nc, dtype=torch.int64, device=dev)
        oc = torch.empty(nc, dtype=torch.int32, device=dev)
        ctx._check(L.rbffd_jittered_lattice_box_device(ctx._h, dim, G, 0, (C.c_int64 * 3)(*(lo + [0] * (3 - dim))), (C.c_int64 * 3)(*(hi + [0] * (3 - dim))),
                                                       (C.c_int32 * 3)(*(list(blocks) + [1] * (3 - dim))), Xc.data_ptr(), gc.data_ptr(), oc.data_ptr()))
        Xh = Xc.cpu().pin_memory()
        ch = torch.empty((nc, n), dtype=torch.int64).pin_memory()
        vh = torch.empty((r, nc, n), dtype=torch.float64).pin_memory()

        def e2e_step():
            ctx._check(L.rbffd_generate_operator_host(ctx._h, byref(opts), c_void_p(Xh.data_ptr()), nc, None, nc, None,
                                                      c_void_p(ch.data_ptr()), c_void_p(vh.data_ptr())))
        h2d = nc * dim * 8
        d2h = nc * n * 8 + r * nc * n * 8
        e2e_call = ("rbffd_generate_operator_host on the rank's block + candidate margin (%d nodes for %d owned rows; pinned host X in, "
                    "int64 colind + fp64 values out); owned rows verified against the shard's rows" % (nc, M))
    e2e_step()
    e2e_step()
    barrier()
    # the call blocks until the caller's host buffers hold the result, so wall clock around it is the end-to-end time;
    # no collective inside the timed region (weight generation needs none), the max over ranks is taken below
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / Ke
    barrier()
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_nodes / e2e_s
    # the same call with int32 indices in the caller's buffer (a SparseMatrixCSC{Float64,Int32}: no widening pass on the host)
    opts32 = rb.make_options(dim, p, n, deg, CFG["ops"], kernel=args.kernel, index_width=32)
    n_e2e = Xh.shape[0] if world == 1 else nc
    ch32 = torch.empty((n_e2e, n), dtype=torch.int32).pin_memory()

    def e2e32_step():
        ctx._check(L.rbffd_generate_operator_host(ctx._h, byref(opts32), c_void_p(Xh.data_ptr()), n_e2e, None, n_e2e, None,
                                                  c_void_p(ch32.data_ptr()), c_void_p(vh.data_ptr())))
    e2e32_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e32_step()
    e2e32_s = (time.perf_counter() - t0) / Ke
    barrier()
    if world > 1:
        t = torch.tensor([e2e32_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e32_s = float(t.item())
    same32 = bool(torch.equal(ch32.to(torch.int64), ch))
    del ch32
    e2e_rows_ok = None
    if world > 1:
        # the owned rows of the host result, mapped to global ids, are the shard's rows (same stencil sets, same weights)
        mine = torch.nonzero(oc == rank).squeeze(1).cpu()                   # candidate ids of the owned nodes, ascending global id
        gcand = gc.cpu()
        gsh = torch.from_numpy(shard.global_ids())
        order = torch.argsort(gsh[:M])                                      # shard rows in ascending global id
        take = min(M, 200000)
        sel_c, sel_s = mine[:take], order[:take]
        cols_host = gcand[ch[sel_c].reshape(-1)].reshape(take, n)
        shard_cols = gsh[colind.cpu()[sel_s].reshape(-1).long()].reshape(take, n)
        e2e_rows_ok = bool(torch.equal(cols_host, shard_cols)) and bool(torch.equal(vh[0][sel_c], vals[0].cpu()[sel_s]))
        if not e2e_rows_ok:
            raise SystemExit(f"rank {rank}: rows of the host call on block + margin differ from the shard's rows")
        del Xc, gc, oc

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"

    # ---- the other BASELINE configs, in the same line
    configs = {}
    if not args.no_configs and args.config == 2:
        del stencils, colind, vals, u, y
        op.close()
        if world == 1:
            del X, Xh, ch, vh
        torch.cuda.empty_cache()
        if world == 1:
            try:
                configs["configs[2]_full_size"] = full_size_config(rb, ctx, torch, dev, CONFIGS[3], 3162, peak, hbm_peak)
                configs["configs[3]_full_size"] = full_size_config(rb, ctx, torch, dev, CONFIGS[4], 271, peak, hbm_peak)
            except Exception as e:                    # report, never hide: the headline numbers above stand on their own
                configs["error"] = repr(e)
            # the drop-in call: generate_operator's full tuple (E, Dx, Dy, Dxx, Dyy, Dxy) on the headline node set, host buffers
            try:
                ops6 = ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"]
                o6 = rb.make_options(dim, p, n, deg, ops6)
                g6 = CFG["g"]
                N6 = g6 ** dim
                X6 = torch.from_numpy(rb.nodes.jittered_lattice(dim, g6, 0)).pin_memory()
                c6 = torch.empty((N6, n), dtype=torch.int64).pin_memory()
                v6 = torch.empty((6, N6, n), dtype=torch.float64).pin_memory()
                call6 = lambda: ctx._check(L.rbffd_generate_operator_host(ctx._h, byref(o6), c_void_p(X6.data_ptr()), N6, None, N6, None,
                                                                          c_void_p(c6.data_ptr()), c_void_p(v6.data_ptr())))
                call6()
                t0 = time.perf_counter()
                for _ in range(2):
                    call6()
                dt6 = (time.perf_counter() - t0) / 2
                configs["e2e_reference_tuple"] = {"call": "generate_operator(X, X, p, n, polydeg) -> (E, Dx, Dy, Dxx, Dyy, Dxy) through rbffd_generate_operator_host",
                                                  "nodes": N6, "ops": ops6, "ms_per_call": dt6 * 1e3, "stencils_per_s": N6 / dt6,
                                                  "h2d_bytes": N6 * dim * 8, "d2h_bytes": N6 * n * 8 + 6 * N6 * n * 8,
                                                  "d2h_gbs": (N6 * n * 4 + 6 * N6 * n * 8) / dt6 * 1e-9}
                del X6, c6, v6
            except Exception as e:
                configs["e2e_reference_tuple"] = {"error": repr(e)}
        else:
            if shard is not None:
                shard.close()
                shard = None
            torch.cuda.empty_cache()
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("adv_diff3d_sharded", os.path.join(ROOT, "examples", "adv_diff3d_sharded.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                out5, _ = mod.run(args.cfg5_g, 12, graph=True)
                out5["spmv_halo_frac_of_hbm"] = out5["spmv_halo_gbs_per_gpu"] / hbm_peak
                configs["configs[4]"] = out5
            except Exception as e:
                configs["configs[4]"] = {"error": repr(e)}

    if rank == 0:
        t_w = phase["weights"] / K * 1e-3
        t_s = phase["spmv"] / K * 1e-3
        t_k = phase["knn"] / K * 1e-3
        F = flops_per_stencil(m, r)
        fp64_peak = max(peak["dfma_tflops"], peak["dmma_tflops"])
        ach_w = F * M / t_w * 1e-12
        ach_s = spmv_bytes_per_row(n) * M / (spmv_loop_ms * 1e-3) * 1e-9
        ach_s_step = spmv_bytes_per_row(n) * M / t_s * 1e-9
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("config%d" % args.config, {})
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("configs[1]: 2D Laplacian operator, %d jittered-lattice nodes per GPU, PHS r^5 + deg-3 polynomials, "
                                    "k=30 (m=%d, r=%d)" % (M, m, r)) if args.config == 2 else
                                   ("configs[%d] shape: %dD, %d nodes per GPU, p=%d, polydeg=%d, k=%d (m=%d, r=%d) ops=%s"
                                    % (args.config - 1, dim, M, p, deg, n, m, r, CFG["ops"])),
                       "step": "exact kNN + fused weight solve -> CSR + one SpMV (halo exchange fused into its launch if N>1); nodes resident in HBM",
                       "l2": "every step writes %.0f MB of stencils+operator (> 126 MB L2), so no input survives in L2 between steps" % ((M * n * 4 * 2 + r * M * n * 8) / 1e6),
                       "parallelism": "spatial blocks %s, halo = stencil closure (%s nodes on rank 0), halo exchange: %s"
                                      % ("x".join(str(b) for b in blocks), "0" if world == 1 else str(NL - M),
                                         "none" if world == 1 else ("NCCL send/recv (RBFFD_HALO=nccl)" if use_nccl else "fused into the SpMV launch, NVLink peer stores into CUDA-IPC inboxes")),
                       "global_nodes": total_nodes},
            "phases_ms": {"knn": phase["knn"] / K, "weights": phase["weights"] / K, "spmv(+halo)": phase["spmv"] / K},
            "roofline": {"kernel": "fused weight kernel (assemble + null-space elimination + solve + CSR write), flops by the LU convention (2/3)m^3 + 2m^2 r", "bound": "fp64",
                         "achieved": ach_w, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_w / fp64_peak,
                         "flop_per_stencil": F, "traffic": traffic.get("weights"),
                         "peak_source": "measured live in this run by rbffd_measure_fp64_peak (DFMA %.2f / DMMA %.2f TFLOP/s burst); "
                                        "MEASURED_PEAKS.json has no FP64 entry" % (peak["dfma_tflops"], peak["dmma_tflops"])},
            "roofline_spmv": {"kernel": "spmv_multi_kernel" if world == 1 else "shard_spmv_kernel (halo exchange fused)", "bound": "hbm", "achieved": ach_s, "peak": hbm_peak, "unit": "GB/s",
                              "frac": ach_s / hbm_peak, "bytes_per_row": spmv_bytes_per_row(n), "traffic": traffic.get("spmv"),
                              "ms": spmv_loop_ms, "achieved_in_step": ach_s_step, "frac_in_step": ach_s_step / hbm_peak,
                              "peak_source": hbm_src,
                              "note": "loop of %d back-to-back applications cycling through 3 copies of the operator (each application streams from HBM), max over ranks, halo exchange included when N>1 (per rank: %d rows); "
                                      "the in-step figure also contains the rank-to-rank skew of the preceding weight solve" % (reps, M)},
            "knn": {"queries_per_s": M / t_k, "ms": t_k * 1e3},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": Ke, "call": e2e_call, "owned_rows_verified": e2e_rows_ok,
                    "int32_indices": {"value": total_nodes / e2e32_s, "ms_per_step": e2e32_s * 1e3, "d2h_bytes_per_step": d2h - (d2h // (8 + 8 * r)) * 4,
                                      "same_pattern_as_int64": same32, "note": "opts.index_width = 32"}},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if world > 1:
            line["sharded_parity"] = parity
            line["shard_setup_ms"] = t_shard * 1e3
            line["host_threads_bound_to_gpu_numa_node"] = bound_cpus
        if configs:
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            v, dt, Ns, cores = cpu_sample(args.ref_sample_g)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{Ns}-node sample (g={args.ref_sample_g}) of the same workload, {dt:.2f} s; CPU oracle = "
                                              "C/OpenMP restatement of the reference (no Julia on the box)"}
        emit(line)
    if shard is not None:
        shard.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
