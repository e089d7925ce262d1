#!/bin/bash
# round 2, step aw: source-level ncu capture of ns2_solve (config 3 and config 4 shapes)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ns2_solve -s 2 -c 1 -f -o gpurun_out/r02aw_ns2_solve_cfg3 python bench.py --config 3 --steps 1 --warmup 1 --profile > gpurun_out/r02aw_ncu_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ns2_solve -s 2 -c 1 -f -o gpurun_out/r02aw_ns2_solve_cfg4 python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/r02aw_ncu_cfg4.log 2>&1
ls -la gpurun_out/*.ncu-rep
