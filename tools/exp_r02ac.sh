#!/bin/bash
# round 2, step ac: cfg4 solve kernel at 5 CTAs per SM (right-hand sides in the spare columns of the Phi~ tile)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02ac_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02ac_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 4 five X=1; q 4 four_by_pad RBFFD_NSW_PAD_SMEM=6000; q 4 five X=1; q 3 six X=1; q 2 base X=1; } | tee gpurun_out/r02ac_sweep.txt
