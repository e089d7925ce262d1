#!/usr/bin/env python
"""Per-source-line / per-phase instruction and stall-sample breakdown of one kernel from an ncu report.
usage: ncu_phases.py report.ncu-rep cubin kernel-substring source.cu [N stencils] [top]"""
import collections, csv, io, re, subprocess, sys
rep, cubin, sub, srcf = sys.argv[1:5]
NS = float(sys.argv[5]) if len(sys.argv) > 5 else 1e6
top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
base_name = srcf.split("/")[-1]
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and sub in l][0]
# -gi prints the whole inlining chain before an instruction (innermost first); the phase of an instruction is decided
# by the OUTERMOST line that lies in the kernel's own source file
line_of = {}; chain = []; cur = None; fresh = False
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not fresh: chain = []; fresh = True
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        if fresh:
            own = [ln for f, ln in chain if f == base_name]
            cur = own[-1] if own else cur
            fresh = False
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]; hdr = rows[hi]
ia, iex, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[hi + 1][ia], 16)
agg = collections.Counter(); sam = collections.Counter(); tot = 0; ts = 0
for r in rows[hi + 1:]:
    if len(r) <= iex: continue
    key = line_of.get(int(r[ia], 16) - base, 0) or 0
    v = int(r[iex] or 0); agg[key] += v; tot += v
    sv = int(r[isamp] or 0); sam[key] += sv; ts += sv
src = open(srcf).read().splitlines()
# phases = comment lines starting with "// ---- "
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith("// ---- ")]
marks = [(0, "(prologue)")] + marks + [(10**9, "")]
print(f"total warp-instructions per stencil: {tot / NS:.0f}")
for (lo, name), (hi2, _) in zip(marks[:-1], marks[1:]):
    v = sum(c for l, c in agg.items() if lo <= l < hi2); s = sum(c for l, c in sam.items() if lo <= l < hi2)
    if v: print(f"  {v / NS:7.0f} instr ({100 * v / tot:4.1f}%)  samples {100 * s / ts:4.1f}%  {name[:90]}")
print("--- top lines")
for l, c in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    txt = src[l - 1].strip()[:80] if 0 < l <= len(src) else ""
    print(f"  line {l:<4d} {c / NS:7.0f} instr  samples {100 * sam[l] / ts:4.1f}%  {txt}")
