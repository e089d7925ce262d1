#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one --set full capture per hot kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
python tools/e2e_trace.py > gpurun_out/e2e_trace.txt 2>&1
python tools/oversampled_bench.py 2 1000 3 > gpurun_out/oversampled_2d.json 2>> gpurun_out/bench_n1.err
python tools/oversampled_bench.py 3 100 2 > gpurun_out/oversampled_3d.json 2>> gpurun_out/bench_n1.err
for c in 3 4; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$c.json 2>> gpurun_out/bench_n1.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_launch.log 2>&1
for k in weights spmv_multi; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_$k.log 2>&1
done
# the k = 30 search is the first knn_kernel launch of a step (the k = 1 search follows it): skip the warm-up step's two
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 2 -c 1 -f -o gpurun_out/prof_knn_kernel \
      python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_knn_kernel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:weights -s 1 -c 1 -f -o gpurun_out/prof_weights_cfg3 \
      python bench.py --config 3 --steps 1 --warmup 1 --profile > gpurun_out/ncu_weights_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:weights -s 1 -c 1 -f -o gpurun_out/prof_weights_cfg4 \
      python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/ncu_weights_cfg4.log 2>&1
ls -la gpurun_out
