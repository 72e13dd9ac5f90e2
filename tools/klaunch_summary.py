#!/usr/bin/env python
"""Per-kernel mean of the metrics in an ncu --csv log (tools/gpu_check.sh klaunch)."""
import csv, sys, re
from collections import OrderedDict
rows=[l for l in open(sys.argv[1]) if l.startswith('"')]
agg=OrderedDict()
for r in csv.DictReader(rows):
    name=re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","")[:40]
    a=agg.setdefault(name,{})
    a.setdefault(r["Metric Name"],[]).append(float(r["Metric Value"].replace(",","")))
for k,m in agg.items():
    print(f"{k:42s}", "  ".join(f"{n.split('.')[0][-28:]}={sum(v)/len(v):10.2f}" for n,v in m.items()), f" n={len(next(iter(m.values())))}")
