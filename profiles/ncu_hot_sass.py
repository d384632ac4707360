#!/usr/bin/env python
"""Hottest SASS instructions of one kernel by warp-stall samples, from `ncu -i rep --page source --csv --kernel-name regex:K`.
usage: python profiles/ncu_hot_sass.py src.csv [top_n]   (shows each hot instruction with 2 instructions of context before it)"""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr_i]
    col = {h: i for i, h in enumerate(H)}
    body = []
    for r in rows[hdr_i + 1:]:          # first launch only (the page repeats per profiled launch)
        if r and r[0] in ("Kernel Name", "Address"):
            break
        if len(r) == len(H):
            body.append(r)
    tot = sum(int(r[col["# Samples"]] or 0) for r in body)
    stall_cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
    print(f"total samples {tot}; top {top} instructions:")
    for i in sorted(idx):
        r = body[i]
        n = int(r[col["# Samples"]] or 0)
        st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        ctx = " | ".join(body[j][col["Source"]].strip()[:40] for j in range(max(0, i - 2), i))
        print(f"{100.0 * n / tot:5.1f}%  #{i:5d} {r[col['Source']].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}   <- {ctx}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
