#!/usr/bin/env python
"""One-screen summary of an ncu --set full report: duration, pipes, stalls, occupancy, DRAM traffic.
usage: ncu_summary.py report.ncu-rep [units-per-launch]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
for r in rows[2:]:
    d = {n: (r[i], u[i]) for i, n in enumerate(h)}
    g = lambda k: d.get(k, ("nan", ""))[0]
    f = lambda k: float(g(k).replace(",", "")) if g(k) not in ("", "nan") else float("nan")
    print("##", g("Kernel Name"), " grid", g("Grid Size"), "block", g("Block Size"))
    dur = f("gpu__time_duration.sum"); du = d["gpu__time_duration.sum"][1]
    print(f"duration {dur} {du}; regs {g('launch__registers_per_thread')}; dyn smem/CTA {g('launch__shared_mem_per_block_dynamic')} KB; "
          f"occupancy limits (CTAs/SM): regs {g('launch__occupancy_limit_registers')} smem {g('launch__occupancy_limit_shared_mem')} warps {g('launch__occupancy_limit_warps')}; "
          f"achieved warps/SM {f('sm__warps_active.avg.per_cycle_active'):.1f}")
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
    print(f"dram read {rd} {d['dram__bytes_read.sum'][1]} + write {wr} {d['dram__bytes_write.sum'][1]}; dram throughput {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} % of ncu peak")
    inst = f("smsp__inst_executed.sum")
    print(f"warp-instructions {inst:.4g}" + (f" = {inst / units:.0f} per unit" if units else "") +
          f"; issue slots busy {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} %; IPC/SM {g('sm__inst_executed.avg.per_cycle_active')}")
    print("pipes (% of peak, active): " + ", ".join(
        f"{k.split('pipe_')[1].split('.')[0]} {float(v[0]):.1f}" for k, v in sorted(d.items())
        if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active") and v[0] not in ("", "0") and float(v[0]) >= 1.0))
    print(f"fp64 tensor path (DMMA) {g('sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed')} % of 128 flop/clk/SM")
    st = [(float(v[0]), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k, v in d.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[0] not in ("",)]
    st.sort(reverse=True)
    print("stalls (warps stalled per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
    print()
