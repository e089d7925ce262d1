#!/bin/bash
# round 2, step u: two GPUs -- shard check incl. the collective transport, bench with either halo transport
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tests/mgpu_shard_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6 | tee gpurun_out/r02u_mgpu_shard_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02u_bench_n2.json 2> gpurun_out/r02u_bench_n2.err
RBFFD_HALO=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 2 --steps 10 --warmup 3 --no-configs > gpurun_out/r02u_bench_n2_nccl.json 2> gpurun_out/r02u_bench_n2_nccl.err
tail -c 600 gpurun_out/r02u_bench_n2.err; tail -c 600 gpurun_out/r02u_bench_n2_nccl.err
python - <<'PY'
import json
for f in ("r02u_bench_n2.json", "r02u_bench_n2_nccl.json"):
    d = json.load(open("gpurun_out/" + f))
    print(f, {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "sharded_parity")}, d["roofline_spmv"]["frac"], d["roofline_spmv"]["frac_in_step"], d["e2e"]["ms_per_step"], d["e2e"]["int32_indices"]["ms_per_step"])
    if "configs" in d: print({k: v for k, v in d["configs"]["configs[4]"].items() if k in ("ms_per_step", "spmv_halo_frac_of_hbm", "generation_s", "rel_l2_error_vs_exact")})
PY
