#!/usr/bin/env python
"""Turn ncu output into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches <launches.csv>          per-kernel totals / shares of a launch list
  python tools/summarize_ncu.py full <report.ncu-rep>            key metrics of every launch in a --set full report
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_pipe_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v * 1e6 if unit == "s" else v
        tot[row["Kernel Name"]][0] += 1
        tot[row["Kernel Name"]][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e3:.3f} ms of kernel time (ncu: serialised, cold cache)")
    print("| share | total ms | launches | kernel |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {100 * v[1] / total:5.1f} % | {v[1] / 1e3:8.3f} | {v[0]} | `{k[:120]}` |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(data)} launches (ncu --set full --clock-control none)")
    print("| kernel | " + " | ".join(n for m, n in METRICS if m in idx) + " |")
    print("|---|" + "---|" * sum(1 for m, _ in METRICS if m in idx))
    for r in data:
        cells = []
        for m, _ in METRICS:
            if m in idx:
                cells.append(f"{r[idx[m]]} {units[idx[m]]}".strip())
        print(f"| `{r[idx['Kernel Name']][:70]}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
