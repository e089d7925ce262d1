#!/bin/bash
# round 2, step h: two GPUs -- multi-GPU pytest cases, sharded bench with its parity check, config 5 shape
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -x -q -k "two_gpus or fused_cons or time_stepping" 2>&1 | tail -8 | tee gpurun_out/r02h_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
tail -c 1500 gpurun_out/r02h_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02h_bench_ref_n2.json 2>/dev/null
