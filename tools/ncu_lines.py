#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page by CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
Needs -lineinfo at compile time.  Prints, per source line: stall samples, warp instructions executed, shared wavefronts.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cubin, sub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate function
    start = None
    for i, l in enumerate(dis):
        if l.startswith(".text.") and sub in l:
            start = i
            break
    assert start is not None, "kernel not found in cubin"
    line_of = {}
    cur = None
    for l in dis[start + 1:]:
        if l.startswith("\t.section") or (l.startswith(".text.") and sub not in l):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    ia, isamp, iex, iwf = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
    isrc = hdr.index("Source")
    base = int(rows[hi + 1][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    opagg = collections.defaultdict(lambda: [0, 0])
    tot = [0, 0, 0]
    for r in rows[hi + 1:]:
        if len(r) <= iwf:
            continue
        off = int(r[ia], 16) - base
        key = line_of.get(off, ("?", 0))
        v = (int(r[isamp] or 0), int(r[iex] or 0), int(r[iwf] or 0))
        for k in range(3):
            agg[key][k] += v[k]
            tot[k] += v[k]
        op = r[isrc].split()[0] if r[isrc].split() else "?"
        if op.startswith("@"):
            op = r[isrc].split()[1]
        op = op.split(".")[0]
        opagg[op][0] += v[0]
        opagg[op][1] += v[1]
    print(f"total samples {tot[0]}  warp-instructions {tot[1]}  shared wavefronts {tot[2]}")
    print("--- by source line (sorted by stall samples)")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0]}:{key[1]:<5d} samples {100 * v[0] / tot[0]:5.1f}%  instr {100 * v[1] / tot[1]:5.1f}%  smem-wf {100 * v[2] / max(tot[2], 1):5.1f}%")
    print("--- by opcode (sorted by instructions)")
    for op, v in sorted(opagg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{op:12s} instr {100 * v[1] / tot[1]:5.1f}%  samples {100 * v[0] / tot[0]:5.1f}%")


if __name__ == "__main__":
    main()
