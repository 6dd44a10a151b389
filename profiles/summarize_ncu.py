#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python profiles/summarize_ncu.py raw gpurun_out/prof.ncu-rep       > profiles/rNN_prof.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, order, tot = collections.OrderedDict(), [], 0.0
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1000 if unit.startswith("n") else (v * 1000 if unit.startswith("m") else v)
        name = row["Kernel Name"][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total "
          "(ncu per-launch times: cold cache, serialised - compare SHARES)")
    print(f"{'us':>12} {'share':>7} {'n':>5}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} {100 * t / tot:6.1f}% {n:5d}  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full, per launch")
    for row in rows[2:]:
        print(f"\n## {row[hdr.index('Kernel Name')][:100]}")
        for m in RAW_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:75s} {row[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
