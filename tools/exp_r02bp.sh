#!/bin/bash
# round 2, step bp: N = 2 bench line repeated (the r02bn line showed every HOST-side figure 3-10x slow: shard set-up, generation, wiring)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02bp_bench_n2.json 2> gpurun_out/r02bp_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r02bp_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'], d.get('shard_setup_ms'))
c=d['configs']['configs[4]']; print(c['generation_s'], c['halo_wiring_s'], c['ms_per_step'], c['spmv_halo_frac_of_hbm'])
"
nproc; uptime
