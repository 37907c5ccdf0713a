#!/usr/bin/env python3
"""ncu `--set full` raw CSV page of consecutive tcgen05 GEMM launches captured in situ -> profiles/*_traffic.json
(per-launch duration, DRAM bytes, tensor-pipe activity; bench.py reads `avg_dram_bytes_per_gemm_launch` for
`roofline.traffic`).
usage: make_insitu_traffic.py raw.csv out.json "source description" """
import csv
import json
import re
import sys

raw, out, source = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, key, to_bytes=False):
    v = float(r[idx[key]].replace(",", ""))
    if to_bytes:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[key]]]
    return v


launches = []
for r in data:
    name = re.sub(r"\(.*$", "", r[idx["Kernel Name"]]).replace("void ", "").replace("gn::", "").replace("<unnamed>::", "")
    us = val(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
                                            "msecond": 1e3}[units[idx["gpu__time_duration.sum"]]]
    launches.append({"kernel": name, "us": us,
                     "dram_read_bytes": val(r, "dram__bytes_read.sum", True),
                     "dram_write_bytes": val(r, "dram__bytes_write.sum", True),
                     "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")})
avg = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches) / len(launches)
json.dump({"source": source, "launches": launches, "avg_dram_bytes_per_gemm_launch": avg}, open(out, "w"), indent=1)
print(len(launches), "launches, avg DRAM bytes per launch", avg)
