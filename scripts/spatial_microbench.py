#!/usr/bin/env python3
"""Timing of the spatial attention kernels through the C ABI test hook (CUDA events; QKV of `frames` frames is
far larger than what one launch leaves in L2 only for big `frames`, so both warm and L2-flushed numbers are printed).
usage: spatial_microbench.py [frames=128] [heads=8] [reps=20] [kernels=0,1]"""
import ctypes as C
import importlib
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("1xgpt_b200")
L = pkg._lib.load()
P = lambda t: C.c_void_p(t.data_ptr())

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
kernels = [int(k) for k in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1]
S, hd = 256, 64
d = H * hd
qkv = torch.randn(frames * S, 3 * d, device="cuda").bfloat16()
out = torch.empty(frames * S, d, device="cuda", dtype=torch.bfloat16)
junk = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
s = torch.cuda.current_stream()
for kernel in kernels:
    for flush in (False, True):
        ts = []
        for i in range(reps + 3):
            if flush:
                junk.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            pkg._lib.check(L.gn_spatial_attention(P(qkv), P(out), frames, S, H, hd, 1.0 / math.sqrt(hd), kernel,
                                                  C.c_void_p(s.cuda_stream)))
            e1.record(s)
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        us = ts[len(ts) // 2] * 1e3
        rows = frames * S
        print(json.dumps({"kernel": {0: "tcgen05", 1: "mma.sync", 2: "generic"}[kernel], "frames": frames, "heads": H,
                          "l2_flush": flush, "us": round(us, 1),
                          "tflops": round(4.0 * S * d * rows / us / 1e6, 1),
                          "GBps_compulsory": round(rows * d * 2 * 4 / us / 1e3, 1)}), flush=True)
