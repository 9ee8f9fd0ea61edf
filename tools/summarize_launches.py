#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and write a markdown
summary (committed under profiles/).  Usage: summarize_launches.py <launches.csv> <out.md> [title]"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"void |vsr::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name


def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else src
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), ns, r.get("Grid Size", "")))
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for k, ns, _ in rows:
        a = agg[k]
        a[0] += 1
        a[1] += ns
        a[2] = max(a[2], ns)
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n{len(rows)} launches, {total / 1e6:.2f} ms of kernel time (ncu-serialised, cold-cache: compare shares)\n\n")
        f.write("| kernel | launches | total ms | share | avg us | max us |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.1f}% | {a[1] / a[0] / 1e3:.1f} | {a[2] / 1e3:.1f} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
