#!/usr/bin/env python3
"""One bench step (GENIE_138M, 64 clips, 8 prompt + 8 generated frames, MaskGIT-2, K/V-cached decode) with eager
launches, for `ncu` launch lists / captures of the production kernels in situ:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file launches.csv \
        python scripts/one_step.py
usage: one_step.py [batch=64] [steps=1]"""
import ctypes as C
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (MODEL_KW, T_PROMPT, MASKGIT_STEPS, synth_state_dict)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = importlib.import_module("1xgpt_b200")
lib = pkg._lib.load()
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(dev))
cfg, sd = bench.synth_state_dict()
m = pkg.STMaskGIT(cfg, precision=os.environ.get("GENIE_PRECISION", "fp16"), kv_cache=True, cuda_graphs=False)
m.load_state_dict(sd)
m = m.to(dev)
h = m._handle()
g = torch.Generator().manual_seed(1234)
T, S = cfg.T, cfg.S
n_new = T - bench.T_PROMPT
clips = torch.randint(0, cfg.image_vocab_size, (B, T, S), generator=g, dtype=torch.int32).to(dev)
noise = torch.stack([torch.stack([torch.stack([torch.randperm(S, generator=g).float() / S for _ in range(B)])
                                  for _ in range(bench.MASKGIT_STEPS - 1)]) for _ in range(n_new)]).to(dev).contiguous()
sptr = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for _ in range(steps):
    work = clips.clone()
    pkg._lib.check(lib.gn_generate(h.ptr, C.c_void_p(work.data_ptr()), B, bench.T_PROMPT, bench.MASKGIT_STEPS, 0.0, 0,
                                   C.c_void_p(noise.data_ptr()), None, None, sptr))
torch.cuda.synchronize()
print("launches", lib.gn_kernel_launches())
