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


def traffic(path, kernel_substr, bench_key, rows_per_launch, out_json):
    """dram bytes (read + write) per launch of the kernels whose name contains `kernel_substr`, averaged over
    the report's launches of it, recorded per input row under `bench_key` in `out_json`
    (profiles/ncu_traffic.json: bench.py's roofline.traffic reads it and scales it to the run's rows per launch).
        python profiles/summarize_ncu.py traffic gpurun_out/prof.ncu-rep refiner_fused gemm_f16x3 614400 profiles/ncu_traffic.json"""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for row in rows[2:]:
        if kernel_substr in row[ik]:
            tot += float(row[ir].replace(",", "")) * scale[units[ir]] + float(row[iw].replace(",", "")) * scale[units[iw]]
            n += 1
    rec = {}
    if os.path.exists(out_json):
        rec = json.load(open(out_json))
    rec[bench_key] = {"dram_bytes_per_row": tot / n / float(rows_per_launch), "dram_bytes_per_launch": tot / n,
                      "rows_per_launch": int(rows_per_launch), "launches_averaged": n, "kernel": kernel_substr,
                      "capture": os.path.basename(path)}
    json.dump(rec, open(out_json, "w"), indent=1)
    print(json.dumps(rec[bench_key]))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:7])
    else:
        {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
