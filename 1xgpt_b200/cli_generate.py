"""Command-line twin of genie/generate.py (same flags, same output files), running the B200 path.

    python -m 1xgpt_b200.cli_generate --checkpoint_dir DIR [--val_data_dir data/val_v1.1] [--output_dir ...]
                                       [--num_prompt_frames 8] [--window_size 16] [--example_ind 0]
                                       [--maskgit_steps 2] [--temperature 0]
"""
import argparse

import torch

from .data import RawTokenDataset
from .generate import generate_and_decode, generate_clips, write_reference_format
from .model import STMaskGIT

STRIDE = 15


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Generates samples (as tokens) from a GENIE model on the B200 path.")
    p.add_argument("--val_data_dir", type=str, default="data/val_v1.1")
    p.add_argument("--checkpoint_dir", type=str, required=True)
    p.add_argument("--output_dir", type=str, default="data/genie_generated")
    p.add_argument("--num_prompt_frames", type=int, default=8)
    p.add_argument("--window_size", type=int, default=16)
    p.add_argument("--example_ind", type=int, default=0)
    p.add_argument("--teacher_force_time", action="store_true")
    p.add_argument("--maskgit_steps", type=int, default=2)
    p.add_argument("--temperature", type=float, default=0)
    p.add_argument("--tokenizer_ckpt", type=str, default=None,
                   help="B200-path extra: a MAGVIT2 Lightning checkpoint (data/magvit2.ckpt); the generated frames are "
                        "decoded on the GPU right after sampling and written to <output_dir>/frames_u8.bin "
                        "(uint8 [new_frames, 3, 16s, 16s])")
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "bf16", "tf32", "fp32"],
                   help="B200-path extra: operand format of the tensor-core kernels (fp16 = parity mode at full speed, "
                        "fp32 = CUDA-core exact mode)")
    return p.parse_args(argv)


@torch.no_grad()
def main(argv=None):
    args = parse_args(argv)
    assert args.num_prompt_frames <= args.window_size
    if args.teacher_force_time:
        raise NotImplementedError("--teacher_force_time crashes in the reference (generate.py:86 uses a non-existent "
                                  "attribute); use the evaluate CLI for temporally teacher-forced decoding")
    ds = RawTokenDataset(args.val_data_dir, window_size=args.window_size, stride=STRIDE)
    s = ds.metadata["s"]
    example = ds[args.example_ind]["input_ids"].reshape(1, args.window_size, s, s)
    model = STMaskGIT.from_pretrained(args.checkpoint_dir, kv_cache=True, precision=args.precision).to("cuda")
    if args.tokenizer_ckpt:
        from .vq import VQModel
        tok = VQModel.from_ckpt(args.tokenizer_ckpt).to("cuda")
        out, imgs = generate_and_decode(model, tok, example, args.num_prompt_frames, args.maskgit_steps, args.temperature)
    else:
        out, imgs = generate_clips(model, example, args.num_prompt_frames, args.maskgit_steps, args.temperature), None
    write_reference_format(args.output_dir, example[0], out[0].cpu(), args.num_prompt_frames, ds.metadata, vars(args))
    if imgs is not None:
        import os
        imgs[0].cpu().numpy().tofile(os.path.join(args.output_dir, "frames_u8.bin"))


if __name__ == "__main__":
    main()
