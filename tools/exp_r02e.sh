#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiwarp or weights_vs_oracle or oversampled or time_stepping or hyperviscosity or mesh_import or chunked" 2>&1 | tail -5 | tee gpurun_out/r02e_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02e_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{
for c in 3 4; do
  q $c fused RBFFD_NS2_SPLIT=0
  q $c split RBFFD_NS2_SPLIT=1
  q $c split_w16 RBFFD_NS2_ELIM_WAVES=16
  q $c split_w256 RBFFD_NS2_ELIM_WAVES=256
  q $c split_occ4 RBFFD_NS2_ELIM_PAD_SMEM=30000
  q $c split_occ6 RBFFD_NS2_ELIM_PAD_SMEM=12000
  q $c split_chunk16k RBFFD_NS2_CHUNK=16384
done
} | tee gpurun_out/r02e_sweep.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02e_launches_cfg4.csv python bench.py --config 4 --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "ns2|knn" gpurun_out/r02e_launches_cfg4.csv | awk -F, '{print $5, $NF}' | tail -7
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02e_launches_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --profile > /dev/null 2>&1
grep -E "ns2|knn" gpurun_out/r02e_launches_cfg3.csv | awk -F, '{print $5, $NF}' | tail -7
