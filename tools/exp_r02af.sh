#!/bin/bash
# round 2, step af: ns2_elim1_kernel (5 x 5 tiles) at 16 warps per SM / 128 registers (spills) vs 12 warps / 168 registers
mkdir -p gpurun_out
P=radialbasisfinitedifferences.jl_b200
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02af_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 3 w12_r168 X=1; q 3 w16_r128 RBFFD_LIB=$P/librbffd_e144.so; q 3 w12_r168 X=1; q 3 w16_r128 RBFFD_LIB=$P/librbffd_e144.so; } | tee gpurun_out/r02af_sweep.txt
