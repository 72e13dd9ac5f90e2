#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean ns, share.

    python tools/summarize_launches.py gpurun_out/launches.csv [skip_launches] > profiles/rNN_launches.md
"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    rows.append((int(r["ID"]), r["Kernel Name"], r["Block Size"], r["Grid Size"], float(r["Metric Value"])))
rows = [r for r in rows if r[0] >= skip]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", "", name) if name.startswith("at::") else name
    return name[:70]


agg = OrderedDict()
for _, name, blk, grd, ns in rows:
    k = (short(name), blk, grd)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print(f"launches {len(rows)} (IDs >= {skip}), total device time {total / 1e3:.1f} us (cold-cache, serialised: compare shares)\n")
print("| kernel | block | grid | launches | mean us | share |")
print("|---|---|---|---:|---:|---:|")
for (name, blk, grd), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {blk} | {grd} | {n} | {ns / n / 1e3:.2f} | {100 * ns / total:.1f}% |")
