#!/usr/bin/env python3
"""Per-shape timing of the tcgen05 linear kernel through the C ABI (CUDA events, L2 flushed or warm)."""
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


def bench(M, N, K, epi, out_bf16, dual, flush, reps=20):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if epi == 2 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if dual else None
    junk = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    s = torch.cuda.current_stream()
    ts = []
    for i in range(reps + 3):
        if flush:
            junk.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        pkg._lib.check(L.gn_linear_forward(P(a), P(w), P(b), P(r if r is not None else None), P(out), P(out2), M, N, K,
                                           epi, 1, int(out_bf16), 0, C.c_void_p(s.cuda_stream)))
        e1.record(s)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    fl = 2.0 * M * N * K
    return med * 1e3, fl / (med * 1e-3) / 1e12


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else None
    Ms = [int(sys.argv[2])] if len(sys.argv) > 2 else (4096, 16384, 65536, 262144)
    shapes = [("qkv", 1536, 512, 0, True, False), ("proj+res(dual)", 512, 512, 2, False, True),
              ("proj+res", 512, 512, 2, False, False), ("fc1+gelu", 2048, 512, 1, True, False),
              ("fc2+res", 512, 2048, 2, False, False), ("readout", 1024, 512, 0, False, False)]
    for M in Ms:
        for name, N, K, epi, obf, dual in shapes:
            if only and only != name:
                continue
            for flush in ((False,) if only else (False, True)):
                us, tf = bench(M, N, K, epi, obf, dual, flush, reps=3 if only else 20)
                print(json.dumps({"M": M, "name": name, "N": N, "K": K, "l2_flush": flush, "us": round(us, 1),
                                  "tflops": round(tf, 1)}), flush=True)
