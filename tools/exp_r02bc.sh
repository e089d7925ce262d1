#!/bin/bash
# round 2, step bc: one launch per Shu-Osher stage (RHS + ghost update + stage combination) on BASELINE config 1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_launch_stage or time_stepping or ghost_node" 2>&1 | tail -5 | tee gpurun_out/r02bc_pytest.log
python - <<'PY' | tee gpurun_out/r02bc_config1_step.txt
import importlib.util, os, sys
spec = importlib.util.spec_from_file_location("adv", "examples/adv_diff_b200.py")
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
for kw in ({"graph": False}, {"graph": True}, {"graph": False, "fused_stage": True}, {"graph": True, "fused_stage": True}):
    t = {}
    mod.run(steps=2000, verbose=False, mesh="tests/golden/rect_0_10.cgns", timing=t, **kw)
    print("config 1 (rect_0_10.cgns, 1812 nodes)", kw, t)
for kw in ({"graph": True}, {"graph": True, "fused_stage": True}):
    t = {}
    mod.run(gy=100, steps=1000, verbose=False, timing=t, **kw)
    print("synthetic rectangle gy = 100", kw, t)
PY
