#!/bin/bash
# round 2, step au: phase clocks of ns2_solve after the symmetric-S change, phase 0 split in three
mkdir -p gpurun_out
rm -f gpurun_out/r02au_timing.txt
for c in 3 4; do
  echo "== NS2_TIMING cfg$c" >> gpurun_out/r02au_timing.txt
  RBFFD_LIB=$PWD/radialbasisfinitedifferences.jl_b200/librbffd_timing.so python bench.py --config $c --steps 1 --warmup 1 --profile 2>&1 | grep "ns2 timing" | tail -8 >> gpurun_out/r02au_timing.txt
done
cat gpurun_out/r02au_timing.txt
