#!/bin/bash
# round 2, step bm: index validation as one vectorised launch, bounding box with one atomic pair per CTA; final checks of the round
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02bm_pytest_gpu.log
{ bash tools/quick_bench.sh 2 3 4; bash tools/quick_bench.sh 2; } | tee gpurun_out/r02bm_quick.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02bm_launches_cfg2.csv python bench.py --config 2 --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "bbox|validate|knn_kernel|weights_ns" gpurun_out/r02bm_launches_cfg2.csv | awk -F'","' '{print $5, $NF}' | tail -8
