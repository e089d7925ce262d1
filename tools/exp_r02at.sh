#!/bin/bash
# round 2, step at: S = V + V' from the basic-node rows of Y only (NS2_SYM) vs the full Y tile, specialised ns2_solve instances
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weights_vs_oracle or nullspace" 2>&1 | tail -3 | tee gpurun_out/r02at_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02at_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
L=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_sym0.so
L4=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_ilp4.so
{ q 3 sym X=1; q 3 full RBFFD_LIB=$L; q 3 sym_ilp4 RBFFD_LIB=$L4; q 4 sym X=1; q 4 full RBFFD_LIB=$L; q 4 sym_ilp4 RBFFD_LIB=$L4; q 3 sym X=1; q 3 full RBFFD_LIB=$L; q 3 sym_ilp4 RBFFD_LIB=$L4; q 4 sym X=1; q 4 full RBFFD_LIB=$L; q 4 sym_ilp4 RBFFD_LIB=$L4; } | tee gpurun_out/r02at_sweep.txt
