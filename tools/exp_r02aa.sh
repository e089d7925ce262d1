#!/bin/bash
# round 2, step aa: two-stage kernels specialised on the configs[2] / configs[3] shapes (n, operator count at compile time)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02aa_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02aa_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ for c in 3 4; do q $c spec3 RBFFD_NS2_SPECIALIZE=3; q $c spec1 RBFFD_NS2_SPECIALIZE=1; q $c generic RBFFD_NS2_SPECIALIZE=0; q $c spec3 RBFFD_NS2_SPECIALIZE=3; done; } | tee gpurun_out/r02aa_sweep.txt
