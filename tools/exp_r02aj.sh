#!/bin/bash
# round 2, step aj: segmented mode of the single-warp kernel (rows of one centre share the elimination)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02aj_pytest.log
{ python tools/oversampled_bench.py 2 1000 3; RBFFD_NS_SEGMENTED=0 python tools/oversampled_bench.py 2 1000 3; python tools/oversampled_bench.py 2 1000 2; RBFFD_NS_SEGMENTED=0 python tools/oversampled_bench.py 2 1000 2; } 2>&1 | tee gpurun_out/r02aj_oversampled.jsonl
