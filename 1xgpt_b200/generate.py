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


@torch.no_grad()
def generate_and_decode(model, tokenizer, clips_THW: torch.Tensor, num_prompt_frames: int = 8, maskgit_steps: int = 2,
                        temperature: float = 0.0, noise: Optional[torch.Tensor] = None, frames: str = "generated",
                        decode_batch: int = 64):
    """Sampling fused with the MAGVIT2 decode (visualize.py:104-120 + eval_utils.py:28-41 without the
    CPU / PIL round trip): gn_generate and gn_vq_decode are enqueued back to back on the current stream of the model's
    device, the tokens never leave the GPU between them and the frames come back as uint8 on the device.

    clips_THW [B,T,H,W]; `tokenizer` is a 1xgpt_b200.VQModel on the same device.  frames = "generated" decodes the
    T - num_prompt_frames new frames, "all" the whole window.
    -> (tokens [B,T,H,W] int64, images uint8 [B, n_frames, 3, 16H, 16W])   (dataset tokens are little-endian)."""
    if tokenizer.device != model.device:
        raise ValueError(f"tokenizer on {tokenizer.device}, model on {model.device}: both must share one GPU / stream")
    out = generate_clips(model, clips_THW, num_prompt_frames, maskgit_steps, temperature, noise)     # [B,T,H,W] on GPU
    B, T, H, W = out.shape
    t0 = 0 if frames == "all" else num_prompt_frames
    flat = out[:, t0:].reshape(-1, H, W)
    imgs = [tokenizer.decode_tokens(flat[s:s + decode_batch], little_endian=True, as_uint8=True)
            for s in range(0, flat.shape[0], decode_batch)]
    imgs = torch.cat(imgs) if len(imgs) > 1 else imgs[0]
    return out, imgs.reshape(B, T - t0, *imgs.shape[1:])


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
