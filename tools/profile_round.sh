#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one --set full capture per hot kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
for c in 3 4; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$c.json 2>> gpurun_out/bench_n1.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_launch.log 2>&1
for k in weights knn_kernel spmv_multi; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:weights -s 1 -c 1 -f -o gpurun_out/prof_weights_cfg4 \
      python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/ncu_weights_cfg4.log 2>&1
ls -la gpurun_out
