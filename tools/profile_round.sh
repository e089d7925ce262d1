#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one --set full capture per hot kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_launch.log 2>&1
for k in weights knn_kernel spmv_multi; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
