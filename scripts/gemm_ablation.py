#!/usr/bin/env python3
"""Where does the tcgen05 linear kernel's time go?  Times one shape with parts of the kernel switched off
(GENIE_B200_GEMM_DEBUG bit 0: epilogue releases the accumulator untouched; bit 1: producer signals stages without
loading; bit 2: epilogue does its math and staging but issues no TMA store; bit 3: epilogue issues its TMA stores
but skips the shared-memory staging writes) for single-CTA and CTA-pair tiles.  Results of the ablated runs are garbage by construction."""
import ctypes as C
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("1xgpt_b200")
L = pkg._lib.load()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())


def run(M, N, K, epi, out_bf16, reps=10):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if epi == 2 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    s = torch.cuda.current_stream()
    ts = []
    for i in range(reps + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        pkg._lib.check(L.gn_linear_forward(P(a), P(w), P(b), P(r), P(out), None, M, N, K, epi, 1, int(out_bf16), 0,
                                           C.c_void_p(s.cuda_stream)))
        e1.record(s)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    return med * 1e3, 2.0 * M * N * K / (med * 1e-3) / 1e12


shapes = [("qkv", 1536, 512, 0, True), ("fc1+gelu", 2048, 512, 1, True), ("fc2+res", 512, 2048, 2, False)]
DBGS = sys.argv[1].split(",") if len(sys.argv) > 1 else ("0", "1", "2", "3")
for pair in ("1", "0"):
    for dbg in DBGS:
        os.environ["GENIE_B200_PAIR"] = pair
        os.environ["GENIE_B200_GEMM_DEBUG"] = dbg
        for name, N, K, epi, obf in shapes:
            for M in (32768, 262144):
                us, tf = run(M, N, K, epi, obf)
                print(json.dumps({"pair": int(pair), "dbg": int(dbg), "name": name, "M": M, "us": round(us, 1),
                                  "tflops": round(tf, 1)}), flush=True)
os.environ["GENIE_B200_GEMM_DEBUG"] = "0"
