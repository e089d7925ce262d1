#!/bin/bash
# round 2, step bl: A/B repeated without the per-call cudaMemGetInfo (total memory read once per process)
mkdir -p gpurun_out
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 8 --warmup 3 --profile 2>gpurun_out/r02bl_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
O=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_old.so
{ for rep in 1 2; do for c in 3 4; do q $c old_build RBFFD_LIB=$O; q $c new_1536MB RBFFD_NS2_SCRATCH_MB=1536; q $c new_12GiB X=1; done; done; } | tee gpurun_out/r02bl_sweep.txt
