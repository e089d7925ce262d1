#!/bin/bash
# round 2, step ay: source-level ncu capture of ns2_elim1 and ns2_pred (config 4 shape), weights_ns (config 2)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ns2_elim1 -s 2 -c 1 -f -o gpurun_out/r02ay_elim1_cfg4 python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/r02ay_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ns2_pred -s 2 -c 1 -f -o gpurun_out/r02ay_pred_cfg4 python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/r02ay_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:weights_ns_kernel -s 1 -c 1 -f -o gpurun_out/r02ay_ns_cfg2 python bench.py --config 2 --steps 1 --warmup 1 --profile > gpurun_out/r02ay_ncu3.log 2>&1
ls -la gpurun_out/r02ay*.ncu-rep
