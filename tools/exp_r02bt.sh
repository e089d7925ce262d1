#!/bin/bash
# round 2, step bt: combined launch geometry (pred waves x solve waves)
mkdir -p gpurun_out
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02bt_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', round(d['phases_ms']['weights'],3))"
}
{ for c in 3 4; do q $c default X=1; q $c pred128_solve64 RBFFD_NS2_PRED_WAVES=128 RBFFD_NSW_WAVES=64; q $c pred512_solve64 RBFFD_NS2_PRED_WAVES=512 RBFFD_NSW_WAVES=64; q $c pred2048_solve64 RBFFD_NS2_PRED_WAVES=2048 RBFFD_NSW_WAVES=64; done; } | tee gpurun_out/r02bt_sweep.txt
