#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals / shares (markdown).
usage: python profiles/summarize.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| `{k[:80]}` | {cnt[k]} | {v:.1f} | {100 * v / T:.1f}% | {v / cnt[k]:.1f} |")
    print(f"\ntotal {T:.1f} us over {sum(cnt.values())} launches (ncu serialises launches with cold caches: compare shares, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1])
