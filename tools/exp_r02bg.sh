#!/bin/bash
# round 2, step bg: ns2_elim1 at 13 / 14 warps per SM (one-warp CTAs, __maxnreg__ 152 / 144, a few spills) vs 12 warps at 168 registers
mkdir -p gpurun_out
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02bg_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
A=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_e1a.so
B=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_e1b.so
{ for c in 3 4; do q $c regs168_12warps X=1; q $c regs152_13warps RBFFD_LIB=$A; q $c regs144_14warps RBFFD_LIB=$B; q $c regs168_12warps X=1; q $c regs152_13warps RBFFD_LIB=$A; done; } | tee gpurun_out/r02bg_sweep.txt
