#!/bin/bash
# Round-2 profile pass, end-of-round build (symmetric S, coordinate layout) (one GPU, under gpurun): ncu launch list of the bench step per config + one --set full capture per hot
# kernel, summarised ON THE BOX (tools/ncu_summary.py); only the summaries and two reports travel back (64 MiB limit).
mkdir -p gpurun_out /tmp/rep
for c in 2 3 4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02bd_launches_cfg$c.csv \
      python bench.py --config $c --steps 2 --warmup 1 --profile > gpurun_out/r02bd_ncu_launch_cfg$c.log 2>&1
done
: > gpurun_out/r02bd_kernels.txt
cap() {  # name regex skip config units-per-launch
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/rep/r02bd_$1 \
      python bench.py --config $4 --steps 1 --warmup 1 --profile > /tmp/rep/ncu_$1.log 2>&1
  echo "=== $1 (bench.py --config $4)" >> gpurun_out/r02bd_kernels.txt
  python tools/ncu_summary.py /tmp/rep/r02bd_$1.ncu-rep $5 >> gpurun_out/r02bd_kernels.txt 2>&1
}
cap weights_ns_cfg2 weights_ns_kernel 1 2 1000000
cap spmv_cfg2 spmv_multi_kernel 1 2 1000000
cap knn_cfg2 knn_kernel 2 2 1000000
cap pred_cfg3 ns2_pred 2 3 86184
cap solve_cfg3 ns2_solve 2 3 86184
cap elim1_cfg3 ns2_elim1 2 3 86184
cap pred_cfg4 ns2_pred 2 4 66576
cap solve_cfg4 ns2_solve 2 4 66576
cap elim1_cfg4 ns2_elim1 2 4 66576
cap knn_cfg4 knn_kernel 2 4 1000000
cp /tmp/rep/r02bd_weights_ns_cfg2.ncu-rep /tmp/rep/r02bd_solve_cfg3.ncu-rep /tmp/rep/r02bd_solve_cfg4.ncu-rep gpurun_out/
tail -n 80 gpurun_out/r02bd_kernels.txt
