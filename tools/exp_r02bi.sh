#!/bin/bash
# round 2, step bi: scratch chunk size of the two-stage weight path (stencils per pred -> solve -> elim1 round)
mkdir -p gpurun_out
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02bi_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ for c in 3 4; do q $c default X=1; q $c chunk150k RBFFD_NS2_CHUNK=150000; q $c chunk260k RBFFD_NS2_CHUNK=260000; q $c chunk520k RBFFD_NS2_CHUNK=520000; q $c chunk1M RBFFD_NS2_CHUNK=1000000; done; } | tee gpurun_out/r02bi_sweep.txt
