"""Command-line twin of the token-space part of genie/evaluate.py (loss / acc / gen_time; the LPIPS leg needs the
`lpips` wheel and `magvit2.ckpt`).  Under torchrun the validation windows are sharded over
the ranks and the metrics are all-reduced once (NCCL).

    [torchrun --nproc-per-node N] python -m 1xgpt_b200.cli_evaluate --checkpoint_dir DIR [--val_data_dir ...]
                                   [--batch_size 16] [--maskgit_steps 2] [--max_examples K]
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

from .data import RawTokenDataset
from .evaluate import b200_backend, evaluate_clips
from .model import STMaskGIT

WINDOW_SIZE, STRIDE = 16, 15


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Evaluate GENIE-style models on the B200 path.")
    p.add_argument("--val_data_dir", type=str, default="data/val_v1.1")
    p.add_argument("--checkpoint_dir", type=str, required=True)
    p.add_argument("--batch_size", type=int, default=16)
    p.add_argument("--maskgit_steps", type=int, default=2)
    p.add_argument("--temperature", type=float, default=0)
    p.add_argument("--max_examples", type=int)
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "bf16", "tf32", "fp32"],
                   help="B200-path extra: operand format of the tensor-core kernels (fp16 = parity mode at full speed, "
                        "fp32 = CUDA-core exact mode)")
    return p.parse_args(argv)


@torch.no_grad()
def main(argv=None):
    args = parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ds = RawTokenDataset(args.val_data_dir, window_size=WINDOW_SIZE, stride=STRIDE, filter_overlaps=True)
    if args.max_examples is not None:
        ds.valid_start_inds = ds.valid_start_inds[:args.max_examples]
    clips = ds.clips()
    model = STMaskGIT.from_pretrained(args.checkpoint_dir, kv_cache=True, precision=args.precision).to(f"cuda:{local}")
    t0 = time.time()
    res = evaluate_clips(clips, b200_backend(model, maskgit_steps=args.maskgit_steps, noise_seed=42,
                                              temperature=args.temperature),
                         batch_size=args.batch_size, acc_device=model.device)
    torch.cuda.synchronize()
    res["gen_time"] = (time.time() - t0) / max(1, (WINDOW_SIZE - 1) * res["local_clips"])
    if rank == 0:
        print(json.dumps({k: (f"{v:.4f}" if isinstance(v, float) else v) for k, v in res.items()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
