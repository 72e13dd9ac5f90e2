#!/usr/bin/env python
"""Read `ncu --page source --csv --print-source sass` output: per kernel, instruction totals, the
hottest SASS instructions by executed count / stall samples, and the stall-reason totals.

    python tools/sass_hot.py gpurun_out/prof_sass.csv [kernel-substring] [topN]
"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
kernels, cur = [], None
with open(path) as f:
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None:
            cur["rows"].append(row)
for k in kernels:
    if filt not in k["name"]:
        continue
    h = {n: i for i, n in enumerate(k["hdr"])}
    rows = k["rows"]
    ie = [int(r[h["Instructions Executed"]] or 0) for r in rows]
    smp = [int(r[h["# Samples"]] or 0) for r in rows]
    print(f"=== {k['name'][:100]}\n  SASS lines {len(rows)}, warp instructions executed {sum(ie):,}, samples {sum(smp):,}")
    stalls = Counter()
    for r in rows:
        for n, i in h.items():
            if n.startswith("stall_") and "Not Issued" not in n:
                stalls[n] += int(r[i] or 0)
    tot = sum(stalls.values()) or 1
    print("  stall mix: " + ", ".join(f"{n[6:]} {100 * v / tot:.1f}%" for n, v in stalls.most_common(8)))
    ops = Counter()
    for r, n in zip(rows, ie):
        ops[r[h["Source"]].split()[0].split(".")[0] if not r[h["Source"]].strip().startswith("@") else r[h["Source"]].split()[1].split(".")[0]] += n
    print("  opcode mix: " + ", ".join(f"{o} {100 * v / max(sum(ie), 1):.1f}%" for o, v in ops.most_common(14)))
    print("  -- hottest by samples")
    for i in sorted(range(len(rows)), key=lambda i: -smp[i])[:topn]:
        r = rows[i]
        print(f"   #{i:5d} smp {smp[i]:6d} exec {ie[i]:9d} thr {r[h['Avg. Threads Executed']]:>4}  {r[h['Source']].strip()[:90]}")
