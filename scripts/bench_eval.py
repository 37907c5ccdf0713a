#!/usr/bin/env python3
"""Teacher-forced evaluation throughput (BASELINE.json configs[2], per-GPU share: 32 clips, K=2, 15 timesteps):
reports evaluate.py's 'gen_time' (seconds per generated frame, evaluate.py:172-175) and frames/s on one GPU
(under torchrun: per rank, plus the all-reduced CE)."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (model config + synthetic weights)

pkg = importlib.import_module("1xgpt_b200")
ev = importlib.import_module("1xgpt_b200.evaluate")
import torch.distributed as dist  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg, sd = bench.synth_state_dict()
m = pkg.STMaskGIT(pkg.GenieConfig(**bench.MODEL_KW), precision=os.environ.get("GENIE_PRECISION", "fp16"), kv_cache=True, chunk_tokens=32768)
m.load_state_dict(sd)
m = m.to(f"cuda:{local}")
clips = torch.randint(0, cfg.image_vocab_size, (B * world, cfg.T * cfg.S), generator=torch.Generator().manual_seed(5))
backend = ev.b200_backend(m, maskgit_steps=2)
for _ in range(2):
    res = ev.evaluate_clips(clips, backend, batch_size=B, acc_device=m.device)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    res = ev.evaluate_clips(clips, backend, batch_size=B, acc_device=m.device)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
frames = 15 * B * world
if rank == 0:
    print(json.dumps({"workload": f"GENIE_138M evaluate.py teacher-forced CE, {B} clips/GPU x {world} GPU, K=2, kv_cache",
                      "ms_per_batch": ms, "gen_time_s_per_frame": ms / 1e3 / frames, "frames_per_s": frames / (ms / 1e3),
                      "loss": res["loss"], "acc": res["acc"], "tokens": res["tokens"]}))
if world > 1:
    dist.destroy_process_group()
