#!/bin/bash
# round 2, step m: block-pivot elimination inside the single-warp kernel (config 2)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02m_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02m_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 2 blockgj X=1; q 2 blockgj_again X=1; } | tee gpurun_out/r02m_sweep.txt
ncu --set full --clock-control none --import-source on -k regex:weights_ns_kernel -s 2 -c 1 -f -o gpurun_out/r02m_weights_ns python bench.py --config 2 --steps 1 --warmup 1 --profile > gpurun_out/r02m_ncu.log 2>&1
