#!/usr/bin/env python3
"""MAGVIT2 encode -> 16x16 LFQ tokens -> decode throughput on one GPU (BASELINE.json configs[4], per-GPU share:
64 synthetic 256x256 frames).  FLOPs per image from SURVEY.md 8d: encoder 135.8 GFLOP, decoder 186.7 GFLOP."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pkg = importlib.import_module("1xgpt_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
# under torchrun: B images PER GPU (batch sharding, no collective); rank 0 reports the max-over-ranks times
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
m = pkg.VQModel(precision=os.environ.get("GENIE_PRECISION", "fp16"))
# seeded synthetic weights (no checkpoint reachable): fan-in scaled convs, GroupNorm affine near identity
_sd = pkg.synthetic_vq_state_dict(m.state_dict(), seed=31)
m.load_state_dict(_sd)
m = m.to("cuda")
img = (torch.rand(B, 3, 256, 256, generator=torch.Generator().manual_seed(7 + rank)) * 2 - 1).cuda()


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


if world > 1:
    dist.barrier()
ms_e, ids = timed(lambda: m.encode_to_tokens(img))
ms_d, _ = timed(lambda: m.decode_tokens(ids, little_endian=False, as_uint8=True))
if world > 1:
    t = torch.tensor([ms_e, ms_d], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e, ms_d = float(t[0]), float(t[1])
    dist.destroy_process_group()
if rank != 0:
    sys.exit(0)
B = B * world
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1400.0}
pk = peaks["bf16_tflops_sustained"]
print(json.dumps({
    "workload": f"MAGVIT2 encode->16x16 LFQ->decode, {B} synthetic 256x256 frames on {world} GPU(s), "
                f"{m.precision} operands", "n_gpus": world,
    "encode_ms": ms_e, "encode_img_s": B / ms_e * 1e3, "encode_tflops": B * 135.8e9 / (ms_e * 1e-3) / 1e12,
    "decode_ms": ms_d, "decode_img_s": B / ms_d * 1e3, "decode_tflops": B * 186.7e9 / (ms_d * 1e-3) / 1e12,
    "roundtrip_img_s": B / (ms_e + ms_d) * 1e3, "peak_tflops_per_gpu": pk,
    "encode_frac": B * 135.8e9 / (ms_e * 1e-3) / 1e12 / pk / world,
    "decode_frac": B * 186.7e9 / (ms_d * 1e-3) / 1e12 / pk / world}))
