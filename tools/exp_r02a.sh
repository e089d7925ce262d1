#!/bin/bash
# r02a: state check + where does the weight-kernel time go.  Runs on the GPU box under gpurun.
#   1. pytest -m gpu
#   2. weight-phase ms of the three BASELINE shapes (device resident) for: baseline, role rotation off/on,
#      occupancy sweep (1..4 CTAs/SM through padded shared memory) -> loaded vs unloaded per-stencil latency
#   3. per-phase clocks (NSW_TIMING build) at 1 and 4 CTAs/SM
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest.log
tail -3 gpurun_out/r02a_pytest.log
q() {  # config, label, env...
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02a_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{
for c in 3 4; do
  q $c rot0 RBFFD_NSW_ROT=0
  q $c rot1 RBFFD_NSW_ROT=1
  # 1, 2, 3 CTAs/SM: pad the ~48 KB tile to > 114, > 76, > 57 KB
  q $c occ1 RBFFD_NSW_ROT=1 RBFFD_NSW_PAD_SMEM=70000
  q $c occ2 RBFFD_NSW_ROT=1 RBFFD_NSW_PAD_SMEM=30000
  q $c occ3 RBFFD_NSW_ROT=1 RBFFD_NSW_PAD_SMEM=12000
done
q 2 base A=1
q 2 occ1 RBFFD_NS_PAD_SMEM=70000
q 2 occ2 RBFFD_NS_PAD_SMEM=30000
q 2 occ3 RBFFD_NS_PAD_SMEM=8000
} | tee gpurun_out/r02a_sweep.txt
for c in 3 4; do
  for pad in 0 70000; do
    echo "== NSW_TIMING cfg$c pad=$pad" >> gpurun_out/r02a_timing.txt
    RBFFD_LIB=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_timing.so RBFFD_NSW_PAD_SMEM=$pad \
      python bench.py --config $c --steps 1 --warmup 1 --profile 2>&1 | grep "nsw timing" | tail -9 >> gpurun_out/r02a_timing.txt
  done
done
cat gpurun_out/r02a_timing.txt
