#!/bin/bash
# round 2, step y: two GPUs after the wait-timeout change: shard check, example test, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py -m gpu -x -q -k "shard or two_gpus" 2>&1 | tail -4 | tee gpurun_out/r02y_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02y_bench_n2.json 2> gpurun_out/r02y_bench_n2.err
tail -c 400 gpurun_out/r02y_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02y_bench_n2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "phases_ms", "sharded_parity")}, d["roofline_spmv"]["frac"], d["roofline_spmv"]["frac_in_step"])
print({k: v for k, v in d["configs"]["configs[4]"].items() if k in ("ms_per_step", "spmv_halo_frac_of_hbm", "generation_s", "rel_l2_error_vs_exact", "error")})
PY
