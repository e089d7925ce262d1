#!/bin/bash
# round 2, step al: segmented mode (DMMA back substitution, 3 / 4 / 5 tile columns); six operators and one operator per row
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02al_pytest.log
{ for args in "2 1000 3" "2 1000 3 lap" "2 1000 2 lap"; do python tools/oversampled_bench.py $args; RBFFD_NS_SEGMENTED=0 python tools/oversampled_bench.py $args; done; } 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d['M'], len(d['ops']), 'generic %.2f ms, null-space %.2f ms = %.1f M rows/s' % (d['generic_kernel1_weights_ms'], d['rowwise_nullspace_weights_ms'], d['rowwise_nullspace_rows_per_s'] * 1e-6))
" | tee gpurun_out/r02al_oversampled.txt
