#!/usr/bin/env python
"""Aggregate an ncu report's source page: stall reasons and hottest SASS lines per kernel.
usage: python benchmarks/ncu_stalls.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
data = rows[rows.index(hdr) + 1:]
H = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
total = 0
per = []
for r in data:
    try:
        n = int(r[H["# Samples"]])
    except (ValueError, IndexError):
        continue
    total += n
    st = {s: int(r[H[s]] or 0) for s in stalls}
    for s in stalls:
        tot[s] += st[s]
    per.append((n, r[H["Source"]].strip(), st))
print(f"{kern}: {total} samples")
print("  " + ", ".join(f"{k[6:]} {100 * v / max(total, 1):.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
per.sort(key=lambda x: -x[0])
for n, src, st in per[:top]:
    print(f"  {100 * n / total:5.1f}% {max(st, key=st.get)[6:]:18s} {src[:100]}")
