#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiwarp or weights_vs_oracle or oversampled or time_stepping or hyperviscosity or mesh_import or chunked" 2>&1 | tail -15 | tee gpurun_out/r02c_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02c_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{
for c in 3 4; do
  q $c ns2 RBFFD_NS2=1
  q $c ns2_occ1 RBFFD_NSW_PAD_SMEM=70000
  q $c ns2_occ2 RBFFD_NSW_PAD_SMEM=30000
  q $c ns2_pw4 RBFFD_NS2_PRED_WAVES=4
  q $c ns2_w32 RBFFD_NSW_WAVES=32
done
} | tee gpurun_out/r02c_sweep.txt
for c in 3 4; do
  echo "== NS2_TIMING cfg$c" >> gpurun_out/r02c_timing.txt
  RBFFD_LIB=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_timing.so python bench.py --config $c --steps 1 --warmup 1 --profile 2>&1 | grep "ns2 timing" | tail -6 >> gpurun_out/r02c_timing.txt
done
cat gpurun_out/r02c_timing.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches_cfg4.csv python bench.py --config 4 --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "ns2|knn" gpurun_out/r02c_launches_cfg4.csv | awk -F, '{print $5, $NF}' | tail -6
