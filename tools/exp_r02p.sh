#!/bin/bash
mkdir -p gpurun_out
python tools/knn_fullsize.py 271 60 | tee gpurun_out/r02p_knn_full.txt
python tools/knn_fullsize.py 100 60 | tee -a gpurun_out/r02p_knn_full.txt
python tools/knn_fullsize.py 200 60 | tee -a gpurun_out/r02p_knn_full.txt
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 0 -c 1 -f -o gpurun_out/r02p_knn_g271 python tools/knn_fullsize.py 271 60 > gpurun_out/r02p_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 0 -c 1 -f -o gpurun_out/r02p_knn_g100 python tools/knn_fullsize.py 100 60 > gpurun_out/r02p_ncu_b.log 2>&1
