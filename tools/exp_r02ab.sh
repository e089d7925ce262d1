#!/bin/bash
# round 2, step ab: cfg3 solve kernel with a shrunk Phi~ tile at 4 / 5 / 6 CTAs per SM; column-reduction kernel specialised
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02ab_pytest.log
P=radialbasisfinitedifferences.jl_b200
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02ab_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 3 mb6 X=1; q 3 mb5 RBFFD_LIB=$P/librbffd_mb5.so; q 3 mb4 RBFFD_LIB=$P/librbffd_mb4.so; q 3 mb6 X=1; q 3 mb6_nopredspec RBFFD_NS2_SPECIALIZE=3;
  q 4 spec7 X=1; q 4 spec3 RBFFD_NS2_SPECIALIZE=3; q 4 spec7 X=1; } | tee gpurun_out/r02ab_sweep.txt
