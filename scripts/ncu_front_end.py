#!/usr/bin/env python
"""Summarise one `ncu --set full` capture per laser front-end kernel into a markdown file.
usage: python scripts/ncu_front_end.py <out.md> <rep> [<rep> ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = ["# ncu summary — laser front-end kernels (SURVEY §8f rows 3, 1, 2)", "",
             "`ncu --set full --clock-control none`, one launch each on 4096 synthetic 1081-beam scans (bench.py `next_rows`).",
             "All three are fp64-ALU / latency-bound: DRAM throughput is a few percent of peak.", ""]
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, r = rows[0], rows[1], rows[2]
        lines += [f"## {r[hdr.index('Kernel Name')].split('(')[0]}", ""]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"- `{w}` = {r[i]} {units[i]}".rstrip())
        st = sorted(((float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                     if "issue_stalled" in h and "ratio" in h and r[i] not in ("", "n/a")), reverse=True)
        lines.append("- top stall reasons (warps per issue): " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st[:4]))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
