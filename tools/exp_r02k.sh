#!/bin/bash
# round 2, step k: eight GPUs -- fused-halo shard check (2x2x2 blocks, real CUDA IPC), the scaling bench line with its parity
# check, config 5 (100M nodes) inside it, the reference arm under torchrun
mkdir -p gpurun_out
SHARD_G=40 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_shard_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -8 | tee gpurun_out/r02k_mgpu_shard_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02k_bench_n8.json 2> gpurun_out/r02k_bench_n8.err
tail -c 1500 gpurun_out/r02k_bench_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_bench_n8.json"))
print({k: d[k] for k in ("value", "ms_per_step", "phases_ms", "sharded_parity", "shard_setup_ms", "host_threads_bound_to_gpu_numa_node")})
print(d["roofline_spmv"]["frac"], d["e2e"])
print(d.get("configs"))
PY
