#!/bin/bash
# round 2, step l: closed-form hyperviscosity right-hand sides in ns2_solve_kernel (config 3 shape, config 1)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02l_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02l_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 3 hvfast X=1; q 4 base X=1; } | tee gpurun_out/r02l_sweep.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02l_launches_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "ns2|knn" gpurun_out/r02l_launches_cfg3.csv | awk -F, '{print $5, $NF}' | tail -4
ncu --set full --clock-control none --import-source on -k regex:ns2_solve -s 2 -c 1 -f -o gpurun_out/r02l_ns2_solve_cfg3 python bench.py --config 3 --steps 1 --warmup 1 --profile > gpurun_out/r02l_ncu_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 1 -c 1 -f -o gpurun_out/r02l_knn_cfg4 python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/r02l_ncu_knn.log 2>&1
