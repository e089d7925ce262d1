#!/bin/bash
# round 2, step ax: conflict-free coordinate layout of the 3-D pair assembly (ns2_solve)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weights_vs_oracle or nullspace" 2>&1 | tail -3 | tee gpurun_out/r02ax_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02ax_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 3 new X=1; q 4 new X=1; q 3 new X=1; q 4 new X=1; } | tee gpurun_out/r02ax_sweep.txt
