#!/bin/bash
# round 2, step bf: batched heap replacement in the kNN kernel (pending lists, warp-convergent flushes) vs the plain kernel
mkdir -p gpurun_out
for pb in 8 16; do
  RBFFD_KNN_BATCH=$pb python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn or boundary_aware or weights_vs_oracle or sorted" 2>&1 | tail -2 | tee -a gpurun_out/r02bf_pytest.log
done
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 5 --warmup 3 --profile 2>gpurun_out/r02bf_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ for c in 2 3 4; do q $c plain X=1; q $c batch8 RBFFD_KNN_BATCH=8; q $c batch16 RBFFD_KNN_BATCH=16; done; } | tee gpurun_out/r02bf_sweep.txt
