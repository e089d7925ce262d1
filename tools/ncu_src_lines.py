#!/usr/bin/env python
"""Per-CUDA-source-line totals of an ncu report captured with --import-source on (-lineinfo build).

usage: ncu_src_lines.py <ncu --page source --csv --print-source cuda,sass output> [top] [units]
Prints the lines with the most stall samples: share of samples, warp instructions (per unit when `units` is given),
shared-memory wavefronts.
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    cur, hdr, out = None, None, []
    with open(path) as f:
        for row in csv.reader(f):
            if len(row) == 2 and row[0] in ("File Path", "File Name"):
                cur = row[1].split("/")[-1]
                continue
            if row and row[0] == "Line No":
                hdr = row
                continue
            if hdr is None or len(row) < 10 or not row[0].isdigit() or row[2] != "-":
                continue
            # duplicated column names ("Source"): index by position
            idx = {name: k for k, name in reversed(list(enumerate(hdr)))}
            try:
                s = int(row[idx["# Samples"]]); i = int(row[idx["Instructions Executed"]]); w = int(row[idx["L1 Wavefronts Shared"]] or 0)
            except ValueError:
                continue
            if s or i:
                out.append((cur, int(row[0]), s, i, w, row[1].strip()[:100]))
    ts = sum(o[2] for o in out) or 1
    ti = sum(o[3] for o in out) or 1
    print("total samples %d, warp instructions %.0f per unit" % (ts, ti / units))
    for o in sorted(out, key=lambda x: -x[2])[:top]:
        print("%-16s %4d  samples %5.1f%%  instr %5.1f%% (%7.1f/unit)  smem wavefronts %7.1f/unit  %s"
              % (o[0], o[1], 100 * o[2] / ts, 100 * o[3] / ti, o[3] / units, o[4] / units, o[5]))


if __name__ == "__main__":
    main()
