#!/bin/bash
# round 2, step i: one-warp block-pivot elimination (ns2_elim1_kernel) vs the two-warp scalar-panel elimination
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02i_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02i_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{
for c in 3 4; do
  q $c elim1 RBFFD_NS2_ELIM=1
  q $c elim2 RBFFD_NS2_ELIM=2
done
q 2 base X=1
} | tee gpurun_out/r02i_sweep.txt
for c in 3 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02i_launches_cfg$c.csv python bench.py --config $c --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "ns2|knn" gpurun_out/r02i_launches_cfg$c.csv | awk -F, '{print $5, $NF}' | tail -4
done
ncu --set full --clock-control none --import-source on -k regex:ns2_elim1 -s 2 -c 1 -f -o gpurun_out/r02i_ns2_elim1 python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/r02i_ncu_elim1.log 2>&1
