#!/usr/bin/env python3
"""A/B of the output conv's activation loads (two 128-bit loads vs one 256-bit load per pixel row) in ONE process:
GENIE_B200_OUT_CONV_L256 is read per launch.  Decoded frames must be bit-identical (same arithmetic, same order); also
checks the default path against the reference-generated fixture like tests/test_gpu_magvit.py does."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import magvit_oracle as MO  # noqa: E402  (checker only)

pkg = importlib.import_module("1xgpt_b200")
z = np.load(os.path.join(ROOT, "tests", "golden", "magvit.npz"), allow_pickle=False)
sd = MO.init_vq_state_dict(MO.VQOracleConfig(), seed=int(z["seed"]))
res = {}
for prec in ("fp16", "bf16"):
    m = pkg.VQModel(precision=prec)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda")
    ids = torch.from_numpy(z["ids"]).long().cuda()
    outs = {}
    for v in ("0", "1"):
        os.environ["GENIE_B200_OUT_CONV_L256"] = v
        outs[v] = (m.decode_tokens(ids, little_endian=False), m.decode_tokens(ids, little_endian=False, as_uint8=True))
    torch.cuda.synchronize()
    ref = torch.from_numpy(z["rec_sub"]).double()
    rel = float(torch.linalg.vector_norm(outs["0"][0][:, :, ::8, ::8].double().cpu() - ref) / torch.linalg.vector_norm(ref))
    res[prec] = {"bit_identical_f32": bool(torch.equal(outs["0"][0], outs["1"][0])),
                 "bit_identical_u8": bool(torch.equal(outs["0"][1], outs["1"][1])), "decode_rel_vs_reference": rel}
# timing on 64 synthetic frames (fp16), 3 reps each, interleaved twice
m = pkg.VQModel(precision="fp16")
m.load_state_dict(pkg.synthetic_vq_state_dict(m.state_dict(), seed=31))
m = m.to("cuda")
ids = torch.randint(0, 262144, (64, 16, 16), generator=torch.Generator().manual_seed(3)).cuda()
times = {"0": [], "1": []}
for rep in range(2):
    for v in ("0", "1"):
        os.environ["GENIE_B200_OUT_CONV_L256"] = v
        m.decode_tokens(ids, little_endian=False, as_uint8=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m.decode_tokens(ids, little_endian=False, as_uint8=True)
        e1.record()
        torch.cuda.synchronize()
        times[v].append(64 * 3 / e0.elapsed_time(e1) * 1e3)
res["decode_img_s_128bit"] = times["0"]
res["decode_img_s_256bit"] = times["1"]
print(json.dumps(res))
