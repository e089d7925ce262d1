#!/bin/bash
# round 2, step bv: the default bench line of the final build (after the launch-geometry change)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02bv_bench_n1.json 2> gpurun_out/r02bv_bench_n1.err
tail -c 300 gpurun_out/r02bv_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r02bv_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline_spmv']['frac'], d['roofline_spmv']['frac_in_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
for k,v in d['configs'].items(): print(k, {kk:v.get(kk) for kk in ('knn_ms','weights_ms','spmv_ms','stencils_per_s','ms_per_call')}, v.get('roofline_weights',{}).get('frac'), v.get('per_pass_ms',{}).get('weights'))
print(d['cpu_baseline'])
"
