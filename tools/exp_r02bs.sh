#!/bin/bash
# round 2, step bs: launch geometry of the two-stage kernels with the larger chunks (stencils per CTA / warp)
mkdir -p gpurun_out
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02bs_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', round(d['phases_ms']['weights'],3))"
}
{ for c in 3 4; do q $c default X=1; q $c solve_w64 RBFFD_NSW_WAVES=64; q $c solve_w16 RBFFD_NSW_WAVES=16; q $c elim_w16 RBFFD_NS2_ELIM_WAVES=16; q $c elim_w256 RBFFD_NS2_ELIM_WAVES=256; q $c pred_w8 RBFFD_NS2_PRED_WAVES=8; q $c pred_w128 RBFFD_NS2_PRED_WAVES=128; done; } | tee gpurun_out/r02bs_sweep.txt
