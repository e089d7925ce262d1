#!/bin/bash
# round 2, step z: single-warp weight kernel specialised on (n = 30, one operator)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02z_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02z_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 2 spec X=1; q 2 generic RBFFD_NS_SPECIALIZE=0; q 2 spec X=1; q 2 generic RBFFD_NS_SPECIALIZE=0; } | tee gpurun_out/r02z_sweep.txt
