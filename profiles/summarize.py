#!/usr/bin/env python
"""Summarise ncu artefacts brought back from the GPU box (run here, no GPU needed).

  python profiles/summarize.py launches gpurun_out/launches_X.csv      # per-kernel mean/min duration
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep          # key metrics + stall breakdown
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    d = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}[row["Metric Unit"]]
        d.setdefault(name, []).append(v)
    tot = sum(sum(v) / len(v) for k, v in d.items())
    print("%-60s %5s %10s %10s" % ("kernel", "n", "mean_us", "min_us"))
    for k, v in d.items():
        print("%-60s %5d %10.1f %10.1f" % (k[:60], len(v), sum(v) / len(v), min(v)))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                print("  %-72s %14s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        items = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
                try:
                    items.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in items) or 1.0
        print("  stalls: " + ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in sorted(items, reverse=True)[:7]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
