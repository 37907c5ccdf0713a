#!/usr/bin/env python3
"""Parity table of the CUDA path against the reference-generated golden fixtures, one JSON line per (fixture, mode):
    python scripts/parity_report.py [--modes fp16 bf16 tf32 fp32] [--out profiles/r02_parity.jsonl]
The numbers printed here are the source of the per-fixture tolerances in tests/test_gpu_model.py (measured x 1.5)."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", nargs="*", default=["fp16", "bf16", "tf32", "fp32"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-long", action="store_true", help="skip the 2-clip teacher-forced eval and the 8-frame generate")
    args = ap.parse_args()
    parity = importlib.import_module("1xgpt_b200.parity")
    lib = importlib.import_module("1xgpt_b200")._lib.load()
    rows = []
    for mode in args.modes:
        for name in ("genie35m", "genie138m", "genie138m_qknorm_mup"):
            f0 = lib.gn_fallback_launches()
            r = parity.production_parity(name, mode, kv_cache=True)
            r["fallback_launches"] = int(lib.gn_fallback_launches() - f0)
            rows.append(r)
            print(json.dumps(r), flush=True)
        if not args.skip_long:
            for fn in (parity.eval_parity, parity.gen8_parity):
                r = fn(mode, kv_cache=True)
                rows.append(r)
                print(json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
