#!/bin/bash
# round 2, step ad: config-2 single-warp kernel in the compact shared-memory layout (5 CTAs per SM)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02ad_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02ad_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 2 compact X=1; q 2 four RBFFD_NS_COMPACT=0; q 2 compact X=1; q 2 four RBFFD_NS_COMPACT=0; q 2 compact_w64 RBFFD_NS_WAVES=64;  q 2 compact_w256 RBFFD_NS_WAVES=256; } | tee gpurun_out/r02ad_sweep.txt
