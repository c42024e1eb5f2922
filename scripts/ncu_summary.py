#!/usr/bin/env python
"""Summarise ncu outputs from gpurun_out/ into profiles/ (tracked).
  launches csv  (ncu --metrics gpu__time_duration.sum --csv --log-file ...)  -> per-kernel count/total/avg/share
  .ncu-rep      (ncu --set full)                                              -> key metrics per captured kernel
usage: python scripts/ncu_summary.py <launches.csv> <prof.ncu-rep> <out.md>"""
import collections
import csv
import subprocess
import sys


def launches(path, full_size_only=False):
    """Per-kernel (count, total ns).  full_size_only: keep, for every kernel, only the launches with that kernel's largest
    grid — bench.py also launches the same kernels on small batches (end-to-end chunks, the single-window probe), which
    would distort the shares of a step of the headline workload."""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    biggest = {}
    if full_size_only:
        for r in rows[hi + 1:]:
            if len(r) > vi:
                name = r[ki].split("(")[0]
                g = int(r[gi].strip("()").split(",")[0])
                biggest[name] = max(biggest.get(name, 0), g)
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        if full_size_only and int(r[gi].strip("()").split(",")[0]) != biggest[r[ki].split("(")[0]]:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] in ("us", "usecond"):
            v *= 1e3
        elif r[ui] in ("ms", "msecond"):
            v *= 1e6
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = collections.OrderedDict()
        d["kernel"] = r[hdr.index("Kernel Name")].split("(")[0]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    return res


def main():
    lcsv, rep, outp = sys.argv[1:4]
    lines = ["# ncu summary", "", f"launch list: `{lcsv}` (cold-cache, serialised: compare shares, not absolutes)", "",
             "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    agg = launches(lcsv)
    tot = sum(a[1] for a in agg.values())
    for k, (c, t) in agg.items():
        lines.append(f"| `{k}` | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {t / tot * 100:.1f}% |")
    lines += ["", "Launches of the headline batch only (each kernel's largest grid; the list above also holds the small-batch launches of",
              "the end-to-end chunks and of the single-window probe):", "", "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    agg = launches(lcsv, full_size_only=True)
    main = {k: v for k, v in agg.items() if any(t in k for t in ("scan_match", "factor_pair", "factor_kernel")) or ("window_kernel" in k and ", 32>" in k)}
    tot = sum(a[1] for a in main.values())
    for k, (c, t) in main.items():
        lines.append(f"| `{k}` | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {t / tot * 100:.1f}% |")
    lines += ["", f"full capture: `{rep}` (`ncu --set full --clock-control none --import-source on`)", ""]
    for d in full(rep):
        lines.append(f"## {d['kernel']}")
        lines.append("")
        for k, v in d.items():
            if k != "kernel":
                lines.append(f"- `{k}` = {v}")
        lines.append("")
    open(outp, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
