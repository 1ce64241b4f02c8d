"""Summaries of ncu output for profiles/.
  python tools/ncu_summarise.py launches <launch-list.csv> <out.csv>     per-kernel totals of a gpu__time_duration launch list
  python tools/ncu_summarise.py full <report.ncu-rep> <out.csv>           selected metrics per launch of a --set full report"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("ia3::", "").replace("at::", "")


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    head = rows[0]
    ik, iv, iu = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    tot, cnt, mx = defaultdict(float), defaultdict(int), defaultdict(float)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        k = short(r[ik])
        tot[k] += v; cnt[k] += 1; mx[k] = max(mx[k], v)
    total = sum(tot.values())
    with open(dst, "w") as fh:
        fh.write("kernel,launches,total_ms,share_pct,max_ms\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            fh.write(f'"{k[:150]}",{cnt[k]},{tot[k]:.3f},{100 * tot[k] / total:.1f},{mx[k]:.3f}\n')
    print(f"{len(rows) - 1} launches, {total:.1f} ms in all -> {dst}")


def full(rep, dst):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    cols = [m for m in METRICS if m in head]
    with open(dst, "w") as fh:
        fh.write("Kernel Name," + ",".join(f"{m} [{units[head.index(m)]}]" for m in cols) + "\n")
        for r in rows[2:]:
            fh.write('"' + short(r[head.index("Kernel Name")]) + '",' + ",".join(r[head.index(m)].replace(",", "") for m in cols) + "\n")
    print(f"{len(rows) - 2} launches -> {dst}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
