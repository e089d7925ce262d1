#!/bin/bash
# round 2, step bb: flat pair index in the Phi~ assembly of the n = 50 instance (ns2_solve)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weights_vs_oracle or nullspace" 2>&1 | tail -3 | tee gpurun_out/r02bb_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02bb_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 3 flat X=1; q 3 flat X=1; q 4 same X=1; } | tee gpurun_out/r02bb_sweep.txt
