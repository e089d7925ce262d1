#!/bin/bash
# round 2, step n: warp-per-query kNN kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02n_pytest.log
RBFFD_KNN_WARP_MIN_K2=8 RBFFD_KNN_WARP_MIN_K3=8 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn or neighbors or boundary_aware or weights" 2>&1 | tail -4 | tee -a gpurun_out/r02n_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02n_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{
q 2 thread X=1
q 2 warp RBFFD_KNN_WARP_MIN_K2=16
q 3 thread RBFFD_KNN_WARP_MIN_K2=100
q 3 warp X=1
q 4 thread RBFFD_KNN_WARP_MIN_K3=100
q 4 warp X=1
q 4 warp_ppc0.9 RBFFD_KNN_PPC=0.9
q 4 warp_ppc1.4 RBFFD_KNN_PPC=1.4
q 3 warp_ppc0.9 RBFFD_KNN_PPC=0.9
q 3 warp_ppc1.4 RBFFD_KNN_PPC=1.4
} | tee gpurun_out/r02n_sweep.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02n_launches_cfg4.csv python bench.py --config 4 --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "knn" gpurun_out/r02n_launches_cfg4.csv | awk -F, '{print $5, $NF}' | tail -6
