#!/usr/bin/env python3
"""Exercises the kernels that the bench step / evaluate script do not launch, for the supplementary ncu capture:
check_masked + logits_transpose (maskgit_generate), relevant_weight + ce with weights (forward), sample_kernel
(temperature > 0), fill_i32 (generate)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pkg = importlib.import_module("1xgpt_b200")
cfg, sd = bench.synth_state_dict()
m = pkg.STMaskGIT(cfg, kv_cache=True)
m.load_state_dict(sd)
m = m.to("cuda")
B = 8
g = torch.Generator().manual_seed(3)
ids = torch.randint(0, cfg.image_vocab_size, (B, cfg.T, 16, 16), generator=g)
for _ in range(2):
    p = ids.clone()
    p[:, 8:] = cfg.image_vocab_size
    m.maskgit_generate(p.cuda(), 8, maskgit_steps=2, temperature=0.0)
    p = ids.clone()
    p[:, 8:] = cfg.image_vocab_size
    m.maskgit_generate(p.cuda(), 8, maskgit_steps=2, temperature=1.0)
    x = ids.clone().reshape(B, -1)
    x[:, 256:][torch.rand(B, 15 * 256, generator=g) < 0.4] = cfg.image_vocab_size
    m(x.cuda(), ids.reshape(B, -1).cuda())
    m.generate(ids[:, :8].reshape(B, -1).cuda(), None, max_new_tokens=8 * 256, maskgit_steps=2)
torch.cuda.synchronize()
print("ok")
