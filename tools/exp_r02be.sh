#!/bin/bash
# round 2, step be: pivot search of the NEXT step of the column reduction issued inside the current step (ns2_pred, weights_ns)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02be_pytest.log
bash tools/quick_bench.sh 2 3 4 | tee gpurun_out/r02be_quick.txt
bash tools/quick_bench.sh 2 3 4 | tee -a gpurun_out/r02be_quick.txt
