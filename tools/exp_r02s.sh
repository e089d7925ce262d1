#!/bin/bash
# round 2, step s: kNN scan restructured (advance-to-next-accepted loop + converged heap update)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02s_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02s_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{ q 2 f32heap X=1; q 3 f32heap X=1; q 4 f32heap X=1; } | tee gpurun_out/r02s_sweep.txt
python tools/knn_fullsize.py 271 60 | tee gpurun_out/r02s_knn_full.txt
