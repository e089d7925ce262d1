#!/bin/bash
# round 2, step bn: two-GPU validation of the final build (scratch budget, fused validation, bounding box)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_shard_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6 | tee gpurun_out/r02bn_mgpu_shard_n2.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02bn_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02bn_bench_n2.json 2> gpurun_out/r02bn_bench_n2.err
tail -c 400 gpurun_out/r02bn_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r02bn_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline_spmv']['frac'], d.get('sharded_parity'), d['e2e']['ms_per_step'])
print(d['configs'])
"
