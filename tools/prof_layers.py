#!/usr/bin/env python
"""Aggregate the per-launch CSV written by the library when CG_PROF_DUMP=<path> is set (bench.py enables the event
profiling of the tensor-core kernels): time and TFLOP/s per distinct launch geometry.
    CG_PROF_DUMP=gpurun_out/tc_launches.csv python bench.py --steps 1 ... ; python tools/prof_layers.py gpurun_out/tc_launches.csv"""
import csv
import sys
from collections import defaultdict

agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in csv.DictReader(open(sys.argv[1])):
    names = {0: "conv", 1: "conv", 2: "conv2cta", 3: "wgrad", 4: "wgrad16", 5: "conv-win", 6: "wgrad-win"}
    key = (names.get(int(r["kind"]), r["kind"]), int(r["taps"]),
           int(r["cchunks"]), int(r["bn"]), int(r["tiles"]), int(r["nb"]))
    a = agg[key]
    a[0] += 1
    a[1] += float(r["flops"])
    a[2] += float(r["ms"])
tot = sum(a[2] for a in agg.values())
print(f"total {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
print("| kind | taps | K-chunks/units | bn | tiles(chunks)/img | images | launches | ms | share | TFLOP/s |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"| {k[0]} | {k[1]} | {k[2]} | {k[3]} | {k[4]} | {k[5]} | {a[0]} | {a[2]:.3f} | {100 * a[2] / tot:.1f}% | {a[1] / a[2] / 1e9 if a[2] else 0:.0f} |")
