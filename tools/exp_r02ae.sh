#!/bin/bash
# round 2, step ae: column reduction of the single-warp kernel publishes the pivot row through shared memory instead of SHFLs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02ae_pytest.log
P=radialbasisfinitedifferences.jl_b200
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02ae_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 2 smem X=1; q 2 shfl RBFFD_LIB=$P/librbffd_old.so; q 2 smem X=1; q 2 shfl RBFFD_LIB=$P/librbffd_old.so; } | tee gpurun_out/r02ae_sweep.txt
