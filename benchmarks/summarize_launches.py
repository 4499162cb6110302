#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and share.
usage: python benchmarks/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if "gpu__time_duration" not in r.get("Metric Name", ""):
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print(f"| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |")
print(f"\ntotal {total / 1e6:.3f} ms over {sum(v[0] for v in tot.values())} launches")
