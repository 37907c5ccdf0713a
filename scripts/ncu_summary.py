#!/usr/bin/env python3
"""ncu `--set full` raw-page CSVs (scripts/r2_ncu.sh) -> one table: per kernel the duration, DRAM bytes, achieved HBM
GB/s and its fraction of the measured copy peak (MEASURED_PEAKS.json), tensor-pipe activity.
usage: ncu_summary.py OUT_PREFIX raw1.csv[.gz] raw2.csv[.gz] ..."""
import csv
import gzip
import io
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def read(path):
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = [r for r in csv.reader(io.StringIO(f.read())) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    return rows[start], rows[start + 1], rows[start + 2:]


def short(name):
    name = re.sub(r"^void ", "", name).replace("gn::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"<?unnamed>::", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    out, files = sys.argv[1], sys.argv[2:]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = float(peaks["hbm_gbs"])
    agg = {}
    for path in files:
        hdr, units, data = read(path)
        idx = {h: i for i, h in enumerate(hdr)}

        def val(r, key, table=None):
            if key not in idx or r[idx[key]] in ("", "n/a"):
                return None
            v = float(r[idx[key]].replace(",", ""))
            return v * table[units[idx[key]]] if table else v

        tkey = next((k for k in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                 "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active") if k in idx), None)
        for r in data:
            if len(r) < len(hdr):
                continue
            k = short(r[idx["Kernel Name"]])
            if k.startswith("at::"):       # torch helper kernels that happened to match a capture regex
                continue
            us = val(r, "gpu__time_duration.sum", TIME)
            rd, wr = val(r, "dram__bytes_read.sum", UNIT) or 0.0, val(r, "dram__bytes_write.sum", UNIT) or 0.0
            a = agg.setdefault(k, {"n": 0, "us": 0.0, "bytes": 0.0, "tensor": 0.0, "dram_pct": 0.0, "src": os.path.basename(path)})
            a["n"] += 1
            a["us"] += us
            a["bytes"] += rd + wr
            a["tensor"] += (val(r, tkey) or 0.0) if tkey else 0.0
            a["dram_pct"] += val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed") or 0.0
    rows = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"] / kv[1]["n"]):
        n = a["n"]
        us, by = a["us"] / n, a["bytes"] / n
        gbs = by / (us * 1e-6) / 1e9 if us > 0 else 0.0
        rows.append({"kernel": k, "launches_captured": n, "avg_us": round(us, 2), "dram_bytes_per_launch": round(by),
                     "hbm_gbs": round(gbs, 1), "hbm_frac_of_measured_copy_peak": round(gbs / hbm, 3),
                     "ncu_dram_throughput_pct": round(a["dram_pct"] / n, 1), "tensor_pipe_active_pct": round(a["tensor"] / n, 1),
                     "capture": a["src"]})
    json.dump({"hbm_peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)",
               "kernels": rows}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("| kernel | launches | avg us | DRAM bytes / launch | HBM GB/s | of copy peak | tensor pipe active % | capture |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---|\n")
        for r in rows:
            f.write(f"| `{r['kernel']}` | {r['launches_captured']} | {r['avg_us']} | {r['dram_bytes_per_launch']:,} | "
                    f"{r['hbm_gbs']} | {r['hbm_frac_of_measured_copy_peak']} | {r['tensor_pipe_active_pct']} | {r['capture']} |\n")
    print(f"{len(rows)} kernels -> {out}.md / .json")


if __name__ == "__main__":
    main()
