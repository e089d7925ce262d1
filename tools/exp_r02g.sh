#!/bin/bash
# round 2, step g: shard path + fused cons_sys + new bench.py on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02g_pytest.log
python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; tail -c 1500 gpurun_out/r02g_bench_n1.err
python - <<'PY' | tee gpurun_out/r02g_config1_step.txt
import importlib.util, os, sys
spec = importlib.util.spec_from_file_location("adv", "examples/adv_diff_b200.py")
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
for kw in ({"graph": False}, {"graph": True}, {"graph": True, "collocated": False}):
    t = {}
    mod.run(steps=400, verbose=False, mesh="tests/golden/rect_0_10.cgns", timing=t, **kw)
    print("config 1 (rect_0_10.cgns, 1812 nodes)", kw, t)
PY
python examples/adv_diff3d_sharded.py --g 100 --steps 30 --graph | tee gpurun_out/r02g_adv3d_n1.json
