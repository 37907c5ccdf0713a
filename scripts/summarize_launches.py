#!/usr/bin/env python3
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel shares as a markdown table.
usage: summarize_launches.py launches.csv [title]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    if len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if "gpu__time_duration" not in r.get("Metric Name", ""):
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name).replace("gn::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "").lstrip("<")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)
    agg[name][0] += 1
    agg[name][1] += ns
total = sum(v[1] for v in agg.values())
print(f"# {title}\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |")
gemm = sum(v[1] for k, v in agg.items() if "gemm_tcgen05" in k)
print(f"\ntotal {total / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches; tcgen05 GEMM variants together "
      f"{100 * gemm / total:.1f}%")
