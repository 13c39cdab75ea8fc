#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  profiles/r1_xxx.csv      # key counters per captured launch
  python tools/ncu_summary.py list gpurun_out/launches.csv  profiles/r1_launches.csv  # per-kernel totals of a launch list
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg",
]


def rep(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEYS]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(open(dst).read())


def launch_list(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mu = hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        if r[mu] == "ns":
            v /= 1e3
        elif r[mu] == "ms":
            v *= 1e3
        t = tot.setdefault(r[kn], [0, 0.0])
        t[0] += 1
        t[1] += v
    total = sum(t[1] for t in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {sum(t[0] for t in tot.values())} launches, "
                f"{total:.1f} us in total (cold-cache, serialised: compare shares)\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{us:.1f},{us / n:.2f},{us / total:.4f}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
