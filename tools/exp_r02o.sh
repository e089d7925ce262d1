#!/bin/bash
# round 2, step o: full GPU suite + the default bench line + reference arm (one GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02o_pytest.log
python bench.py > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err; tail -c 800 gpurun_out/r02o_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02o_bench_ref.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02o_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}); print(d['roofline']['frac'], d['roofline_spmv']['frac'], d['roofline_spmv']['frac_in_step']); print(d['e2e'])
for k,v in d['configs'].items(): print(k, {kk: vv for kk, vv in v.items() if kk in ('knn_ms','weights_ms','spmv_ms','ms_per_call','stencils_per_s','error')}, v.get('roofline_weights',{}).get('frac'), v.get('roofline_spmv',{}).get('frac'))
"
