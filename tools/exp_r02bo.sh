#!/bin/bash
# round 2, step bo: one-shot generation of a 12.5 M-node 3-D shard with the 12 GiB / 1.5 GiB scratch budget (allocation trace on)
mkdir -p gpurun_out
for mb in 12288 1536 12288; do
  echo "== RBFFD_NS2_SCRATCH_MB=$mb"
  RBFFD_TRACE_ALLOC=1 RBFFD_NS2_SCRATCH_MB=$mb python examples/adv_diff3d_sharded.py --g 232 --steps 3 --graph 2>&1 | grep -v "^$" | cut -c1-400 | tail -12
done 2>&1 | tee gpurun_out/r02bo_oneshot.txt
