#!/usr/bin/env python3
"""Condense an ncu report (gpurun_out/*.ncu-rep, read here with `ncu -i`) or a launch list
(`--metrics gpu__time_duration.sum --csv`) into the small text tables kept under profiles/.

    python tools/ncu_summary.py rep    gpurun_out/prof_spmma_r1b.ncu-rep > profiles/r01_spmma_ncu.csv
    python tools/ncu_summary.py launch gpurun_out/launches_r1b.csv      > profiles/r01_launches.csv
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    name_i = head.index("Kernel Name")
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [f"launch{i}:{r[name_i].split('(')[0][-40:]}" for i, r in enumerate(data)])
    for k in KEYS:
        if k in head:
            i = head.index(k)
            w.writerow([k, units[i]] + [r[i] for r in data])


def launch(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    head = rows[0]
    ni, vi = head.index("Kernel Name"), head.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) != len(head):
            continue
        k = r[ni].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_us", "avg_us", "share"])
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, f"{t/1e3:.1f}", f"{t/1e3/n:.2f}", f"{t/tot:.4f}"])


if __name__ == "__main__":
    {"rep": rep, "launch": launch}[sys.argv[1]](sys.argv[2])
