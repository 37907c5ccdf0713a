#!/usr/bin/env python3
"""One MAGVIT2 encode pass + one decode pass of B images (default 32 = one pass through the trunk) for `ncu` launch
lists: the warm-up runs outside the profiled range (ncu --profile-from-start off):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file l.csv \
        python scripts/magvit_one_pass.py 32
Weights / images as in scripts/bench_magvit.py (seeded synthetic)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("1xgpt_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = pkg.VQModel(precision=os.environ.get("GENIE_PRECISION", "fp16"))
sd = pkg.synthetic_vq_state_dict(m.state_dict(), seed=31)
m.load_state_dict(sd)
m = m.to("cuda")
img = (torch.rand(B, 3, 256, 256, generator=torch.Generator().manual_seed(7)) * 2 - 1).cuda()
ids = m.encode_to_tokens(img)
m.decode_tokens(ids, little_endian=False, as_uint8=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ids = m.encode_to_tokens(img)
out = m.decode_tokens(ids, little_endian=False, as_uint8=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("images", B, "launches", pkg._lib.load().gn_kernel_launches())
