#!/bin/bash
# round 2, step bj: two-stage weight path with a 12 GiB scratch budget (larger chunks), stage A zeroing the unmapped positions itself
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02bj_pytest_gpu.log
{ bash tools/quick_bench.sh 2 3 4; bash tools/quick_bench.sh 3 4; for mb in 1536 4096 24576; do echo "RBFFD_NS2_SCRATCH_MB=$mb"; RBFFD_NS2_SCRATCH_MB=$mb bash tools/quick_bench.sh 3 4; done; } | tee gpurun_out/r02bj_quick.txt
timeout 900 python bench.py > gpurun_out/r02bj_bench_n1.json 2> gpurun_out/r02bj_bench_n1.err
tail -c 300 gpurun_out/r02bj_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r02bj_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline_spmv']['frac'], d['e2e']['ms_per_step'])
for k,v in d['configs'].items(): print(k, {kk:v.get(kk) for kk in ('knn_ms','weights_ms','spmv_ms','stencils_per_s','ms_per_call')}, v.get('roofline_weights',{}).get('frac'), v.get('per_pass_ms',{}).get('weights'))
"
