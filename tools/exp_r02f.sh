#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02f_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02f_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{
for c in 3 4; do
  q $c split RBFFD_NS2_SPLIT=1
done
} | tee gpurun_out/r02f_sweep.txt
for c in 3 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches_cfg$c.csv python bench.py --config $c --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "ns2|knn" gpurun_out/r02f_launches_cfg$c.csv | awk -F, '{print $5, $NF}' | tail -4
done
for k in ns2_solve ns2_elim; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02f_$k python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/r02f_ncu_$k.log 2>&1
done
