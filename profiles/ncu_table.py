#!/usr/bin/env python
"""Per-kernel table from an `ncu --set full` report: duration, DRAM bytes and throughput against the measured copy peak,
tensor-pipe activity, registers, achieved occupancy, L2 hit rate.
usage:  ncu -i report.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_table.py raw.csv [peak_GBs] > profiles/ncu_rNN.md
ncu flushes caches between replays (cold data) and serialises launches: the DRAM figures are per launch on cold inputs."""
import csv
import re
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path, peak=6576.7):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    col = {h: i for i, h in enumerate(H)}

    def get(r, name, unit_scale=None):
        i = col.get(name)
        if i is None:
            return float("nan")
        v = num(r[i])
        if unit_scale:
            v *= unit_scale.get(U[i], 1.0)
        return v
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    to_mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    print("| kernel | grid | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | % of copy peak | tensor pipe active % | regs | achieved occ % | L2 hit % |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    seen = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("npf::", "")
        key = (name, r[col["Grid Size"]], round(get(r, "dram__bytes_read.sum", to_mb), 0))
        if key in seen:
            continue
        seen[key] = 1
        us = get(r, "gpu__time_duration.sum", to_us)
        rd, wr = get(r, "dram__bytes_read.sum", to_mb), get(r, "dram__bytes_write.sum", to_mb)
        gbs = (rd + wr) / us * 1e3 if us == us and us > 0 else float("nan")
        tens = get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        if tens != tens:
            tens = get(r, "sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active")
        print(f"| `{name[:56]}` | {r[col['Grid Size']]} | {us:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} | "
              f"{tens:.1f} | {get(r, 'launch__registers_per_thread'):.0f} | {get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{get(r, 'lts__t_sector_hit_rate.pct'):.0f} |")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 6576.7)
