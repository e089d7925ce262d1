#!/bin/bash
# round 2, step bh: eight-GPU bench line of the end-of-round build (configs[4]: 100 M nodes, 2 x 2 x 2 blocks)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02bh_bench_n8.json 2> gpurun_out/r02bh_bench_n8.err
tail -c 300 gpurun_out/r02bh_bench_n8.err
python -c "
import json
d=json.load(open('gpurun_out/r02bh_bench_n8.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline_spmv']['frac'], d.get('sharded_parity'), d['e2e']['ms_per_step'])
print(d['configs'])
"
