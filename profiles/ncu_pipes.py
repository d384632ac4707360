#!/usr/bin/env python
"""Which unit binds?  Per-kernel pipe utilisation table from an `ncu --set full` report (raw csv page):
duration, tensor / XU (MUFU) / FMA / ALU pipe activity, issue-slot utilisation, shared-memory wavefronts, occupancy,
registers and the three largest warp-stall reasons.
usage: ncu -i report.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_pipes.py raw.csv > profiles/ncu_rNN_x.md"""
import csv
import re
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    col = {h: i for i, h in enumerate(H)}
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

    def g(r, name):
        i = col.get(name)
        return num(r[i]) if i is not None else float("nan")
    stall_cols = [h for h in H if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "selected" not in h]
    print("| kernel | grid x block | us | tensor pipe % | XU (MUFU) % | FMA % | ALU % | issue slots % | smem wavefronts % | occupancy % | regs | DRAM GB/s | top stalls (warps per issue) |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    seen = set()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("npf::", "")
        key = (name, r[col["Grid Size"]])
        if key in seen:
            continue
        seen.add(key)
        us = g(r, "gpu__time_duration.sum") * to_us.get(U[col["gpu__time_duration.sum"]], 1.0)
        to_mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        mb = sum(g(r, n) * to_mb.get(U[col[n]], 1.0) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum") if n in col)
        stalls = sorted(((g(r, h), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall_cols), reverse=True)[:3]
        print(f"| `{name[:48]}` | {r[col['Grid Size']]} x {r[col['Block Size']]} | {us:.1f} | {g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f} | {g(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g(r, 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {g(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f} | {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{g(r, 'launch__registers_per_thread'):.0f} | {mb / us * 1e3:.0f} | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
