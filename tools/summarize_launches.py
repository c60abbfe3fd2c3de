#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, val * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    print(f"launches: {len(rows)}   total device time: {tot / 1e3:.2f} ms (cold-cache, serialised under ncu)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / 1e3:.3f} | {100 * t / tot:.1f}% | {t / c:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
