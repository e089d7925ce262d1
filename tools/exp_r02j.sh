#!/bin/bash
# round 2, step j: two GPUs -- sharded bench with the reworked e2e leg; SpMV stage epilogue with early operand loads
mkdir -p gpurun_out
python examples/adv_diff3d_sharded.py --g 100 --steps 30 --graph > gpurun_out/r02j_adv3d_n1.json 2>gpurun_out/r02j_adv3d_n1.err; cat gpurun_out/r02j_adv3d_n1.json | python -c "import json,sys; d=json.load(sys.stdin); print('adv3d n1', d['ms_per_step'], d['spmv_halo_gbs_per_gpu'], d['rel_l2_error_vs_exact'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02j_bench_n2.json 2> gpurun_out/r02j_bench_n2.err
tail -c 1200 gpurun_out/r02j_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_bench_n2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "phases_ms", "sharded_parity", "shard_setup_ms")})
print(d["e2e"])
print(d.get("configs"))
PY
