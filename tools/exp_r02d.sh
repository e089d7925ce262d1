#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiwarp or weights_vs_oracle or oversampled or hyperviscosity" 2>&1 | tail -5 | tee gpurun_out/r02d_pytest.log
q() {
  local c=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $c --steps 3 --warmup 2 --profile 2>gpurun_out/r02d_err.log | python -c "
import json,sys
d=json.load(sys.stdin); print('cfg$c $label', {k: round(v,3) for k,v in d['phases_ms'].items()})"
}
{
for c in 3 4; do
  q $c ns2 RBFFD_NS2=1
  q $c ns2_w32 RBFFD_NSW_WAVES=32
  q $c ns2_w8 RBFFD_NSW_WAVES=8
done
} | tee gpurun_out/r02d_sweep.txt
for k in ns2_pred ns2_solve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02d_$k python bench.py --config 4 --steps 1 --warmup 1 --profile > gpurun_out/r02d_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
