#!/bin/bash
# weight-phase time of the three BASELINE shapes (device resident), one line each
for c in ${@:-2 3 4}; do
  python bench.py --config $c --steps 5 --warmup 3 --profile 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('config', $c, {k: round(v,3) for k,v in d['phases_ms'].items()})"
done
