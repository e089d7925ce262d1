#!/bin/bash
# round 2, step v: the scaling bench at N GPUs (N = $1) + the reference arm under torchrun
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02v_bench_n$N.json 2> gpurun_out/r02v_bench_n$N.err
tail -c 800 gpurun_out/r02v_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02v_bench_ref_n$N.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/r02v_bench_n$N.json"))
print({k: d[k] for k in ("value", "ms_per_step", "phases_ms", "sharded_parity", "shard_setup_ms", "host_threads_bound_to_gpu_numa_node")})
print("spmv loop frac", d["roofline_spmv"]["frac"], "in step", d["roofline_spmv"]["frac_in_step"], "e2e", d["e2e"]["ms_per_step"], "int32", d["e2e"]["int32_indices"]["ms_per_step"])
print({k: v for k, v in d["configs"]["configs[4]"].items() if k in ("global_nodes", "ms_per_step", "spmv_halo_frac_of_hbm", "generation_s", "halo_wiring_s", "rel_l2_error_vs_exact", "error")})
r = json.load(open("gpurun_out/r02v_bench_ref_n$N.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
