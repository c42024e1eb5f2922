#!/usr/bin/env python
"""Per-source-line view of an ncu capture: warp-stall samples and executed warp instructions of every SASS instruction
(ncu --page source --csv) attributed to the CUDA source line nvdisasm -g reports for that offset.
usage: python scripts/ncu_hotlines.py <prof.ncu-rep> <mangled kernel name substring> [top N] [lib.so]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lib = os.path.abspath(sys.argv[4]) if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2dliw-slam_b200", "csrc", "liblvio2d.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, inside = None, False
for ln in dis:
    if ln.startswith("//---------------------"):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
si, ii = h.index("# Samples"), h.index("Instructions Executed")
base = None
samp, inst = collections.Counter(), collections.Counter()
ts = ti = 0
for r in rows[hi + 1:]:
    if len(r) <= ii or not r[ii].isdigit():
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    key = line_of.get(a - base, ("?", 0))
    samp[key] += int(r[si]); inst[key] += int(r[ii])
    ts += int(r[si]); ti += int(r[ii])
print(f"{rep}: {ts} samples, {ti} warp instructions")
src = {}
print("  samples%  inst%   file:line   source")
for key, c in samp.most_common(top):
    f, l = key
    text = ""
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2dliw-slam_b200", "csrc", f)
    if os.path.exists(p):
        if p not in src:
            src[p] = open(p).read().splitlines()
        if 0 < l <= len(src[p]):
            text = src[p][l - 1].strip()[:110]
    print(f"  {c / ts * 100:6.1f}  {inst[key] / ti * 100:6.1f}   {f}:{l}   {text}")
