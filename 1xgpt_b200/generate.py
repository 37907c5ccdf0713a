"""Batched counterpart of genie/generate.py:62-116: prompt frames -> autoregressively generated frames,
written in the reference's on-disk format ([prompt | generated | ground-truth] frames in `video.bin`
+ `metadata.json`).  The reference generates ONE example per invocation (generate.py:70-74); here any number of
clips go through gn_generate in one call, and `example_ind` semantics are kept by `write_reference_format`.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Optional

import numpy as np
import torch


@torch.no_grad()
def generate_clips(model, clips_THW: torch.Tensor, num_prompt_frames: int = 8, maskgit_steps: int = 2,
                   temperature: float = 0.0, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """clips_THW [B,T,H,W] (only the first `num_prompt_frames` frames are read) -> [B,T,H,W] int64 with frames
    >= num_prompt_frames generated (generate.py:77-103 without --teacher_force_time)."""
    B, T, H, W = clips_THW.shape
    cfg = model.config
    assert num_prompt_frames <= T
    prompt = clips_THW[:, :num_prompt_frames].reshape(B, -1)
    out = model.generate(prompt, None, max_new_tokens=(T - num_prompt_frames) * cfg.S, maskgit_steps=maskgit_steps,
                         temperature=temperature, noise=noise)
    return out.reshape(B, T, H, W)


def write_reference_format(output_dir, example_THW: torch.Tensor, generated_THW: torch.Tensor, num_prompt_frames: int,
                           metadata: dict, extra_args: Optional[dict] = None):
    """generate.py:97-116: outputs = [prompt frames, predicted frames, ground-truth frames] of ONE example,
    token_dtype from the dataset metadata, metadata.json = args | dataset metadata | {num_images,h,w,t}."""
    assert example_THW.dim() == 3 and generated_THW.shape == example_THW.shape
    T, H, W = example_THW.shape
    outputs = torch.cat([example_THW[:num_prompt_frames], generated_THW[num_prompt_frames:],
                         example_THW[num_prompt_frames:]], dim=0)
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    dtype = np.dtype(metadata.get("token_dtype", "uint32"))
    outputs.cpu().numpy().astype(dtype).tofile(output_dir / "video.bin")
    meta = dict(extra_args or {})
    meta.update(metadata)
    meta.update({"num_images": int(outputs.shape[0]), "h": H, "w": W, "t": T})
    with open(output_dir / "metadata.json", "w") as f:
        json.dump(meta, f)
    return outputs
