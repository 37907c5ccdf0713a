#!/usr/bin/env python3
"""generate.py throughput for an arbitrary GENIE shape (parity configs of BASELINE.json that are not the headline):
    python scripts/bench_generate.py --layers 40 --d-model 1024 --heads 16 --batch 16 --maskgit-steps 8   # configs[3] per GPU
Random-init weights, synthetic clips, 8 prompt + 8 generated frames, K/V-cached decode, CUDA events."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("1xgpt_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=40)
ap.add_argument("--d-model", type=int, default=1024)
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--maskgit-steps", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--dense", action="store_true")
ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "tf32"])
ap.add_argument("--qk-norm", action="store_true", help="qk_norm=True, use_mup=True (GenieConfig defaults of the reference)")
a = ap.parse_args()

# under torchrun: one replica per GPU (generate has no collective); rank 0 reports the max-over-ranks time
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
cfg = pkg.GenieConfig(num_layers=a.layers, num_heads=a.heads, d_model=a.d_model, T=16, S=256, image_vocab_size=262144,
                      num_factored_vocabs=2, qk_norm=a.qk_norm, use_mup=a.qk_norm)
m = pkg.STMaskGIT(cfg, precision=a.precision, kv_cache=not a.dense)
m.load_state_dict(pkg.synthetic_state_dict(cfg, seed=0, bias_std=0.02))
m = m.to(dev)
h = m._handle()
lib = pkg._lib.load()
B, T, S, K, TP = a.batch, cfg.T, cfg.S, a.maskgit_steps, 8
g = torch.Generator().manual_seed(99 + rank)
clips = torch.randint(0, cfg.image_vocab_size, (B, T, S), generator=g, dtype=torch.int32).to(dev)
noise = torch.rand(T - TP, max(K - 1, 1), B, S, generator=g).to(dev)
sptr = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def step():
    work = clips.clone()
    pkg._lib.check(lib.gn_generate(h.ptr, C.c_void_p(work.data_ptr()), B, TP, K, 0.0, 0, C.c_void_p(noise.data_ptr()),
                                   None, None, sptr))


for _ in range(a.warmup):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
m.reset_counters()
f0 = lib.gn_fallback_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    dist.destroy_process_group()
if rank != 0:
    sys.exit(0)
frames = B * (T - TP) * world
dense_flops = m.flops_per_clip_forward() * B * (T - TP) * K
print(json.dumps({"workload": f"GENIE L{a.layers} d{a.d_model} h{a.heads} generate, {B} clips, MaskGIT-{K}, "
                              f"{'dense' if a.dense else 'K/V-cached'}{', qk_norm+muP' if a.qk_norm else ''}",
                  "params_M": round(sum(p.numel() for p in m.parameters()) / 1e6, 1), "n_gpus": world,
                  "precision": a.precision, "fallback_launches_per_step": int(lib.gn_fallback_launches() - f0) // a.steps,
                  "ms_per_step": ms, "frames_per_s": frames / (ms / 1e3),
                  "executed_tflops_per_gpu": m.flops_executed() / a.steps / (ms / 1e3) / 1e12,
                  "dense_equivalent_tflops_per_gpu": dense_flops / (ms / 1e3) / 1e12}))
