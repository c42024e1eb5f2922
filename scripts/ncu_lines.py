#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py <rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, agg, tot_i, tot_s = "?", {}, 0, 0
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ii or not r[0].isdigit():
        continue
    try:
        n, s = int(r[ii]), int(r[si])
    except ValueError:
        continue
    key = (fname, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0])
    a[0] += n
    a[1] += s
    tot_i += n
    tot_s += s
print(f"kernel {kre}: {tot_i} warp-instructions, {tot_s} samples")
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / tot_i * 100:5.1f}% instr {s / max(tot_s, 1) * 100:5.1f}% samp  {f}:{ln}  {src}")
