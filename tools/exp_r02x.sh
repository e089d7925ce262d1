#!/bin/bash
# round 2, step x: next-pivot-block factorisation software-pipelined inside block_gj_warp
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02x_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 4 --warmup 2 --profile 2>gpurun_out/r02x_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 2 pipelined X=1; q 3 pipelined X=1; q 4 pipelined X=1; } | tee gpurun_out/r02x_sweep.txt
for c in 3 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02x_launches_cfg$c.csv python bench.py --config $c --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "ns2" gpurun_out/r02x_launches_cfg$c.csv | awk -F, '{print $5, $NF}' | tail -3
done
