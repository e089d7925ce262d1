#!/bin/bash
# round 2, step t: knobs of the end-to-end host call (row-chunk pipeline)
mkdir -p gpurun_out
{
echo "default";              python tools/e2e_sweep.py
echo "default int32";        python tools/e2e_sweep.py 32
for c in 2 3 4 6 8; do echo "chunks=$c ratio=0.62"; RBFFD_HOST_CHUNKS=$c python tools/e2e_sweep.py; done
for r in 0.5 0.75 1.0; do echo "chunks=5 ratio=$r"; RBFFD_HOST_CHUNKS=5 RBFFD_HOST_CHUNK_RATIO=$r python tools/e2e_sweep.py; done
echo "chunks=8 ratio=0.8";   RBFFD_HOST_CHUNKS=8 RBFFD_HOST_CHUNK_RATIO=0.8 python tools/e2e_sweep.py
echo "noship";               RBFFD_DEBUG_NOSHIP=1 python tools/e2e_sweep.py
echo "widen threads 4";      RBFFD_WIDEN_THREADS=4 python tools/e2e_sweep.py
echo "widen threads 12";     RBFFD_WIDEN_THREADS=12 python tools/e2e_sweep.py
echo "trace";                RBFFD_TRACE=1 python tools/e2e_sweep.py 2>&1 | tail -12
} 2>&1 | tee gpurun_out/r02t_e2e_sweep.txt
