#!/bin/bash
# round 2, step az: checkpoint after the ns2_solve changes (symmetric S, coordinate layout, dead work of the split path removed):
# full GPU test suite, quick phases of the three BASELINE shapes, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02az_pytest_gpu.log
bash tools/quick_bench.sh 2 3 4 | tee gpurun_out/r02az_quick.txt
timeout 900 python bench.py > gpurun_out/r02az_bench_n1.json 2> gpurun_out/r02az_bench_n1.err
tail -c 300 gpurun_out/r02az_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r02az_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline_spmv']['frac'], d['e2e']['ms_per_step'])
for k,v in d['configs'].items(): print(k, {kk:v.get(kk) for kk in ('knn_ms','weights_ms','spmv_ms','stencils_per_s','ms_per_call')}, v.get('roofline_weights',{}).get('frac'))
"
