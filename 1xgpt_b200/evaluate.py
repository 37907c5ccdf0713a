"""Batch-sharded, multi-GPU version of genie/evaluate.py's metric loop (evaluate.py:146-191).

The reference evaluates on one GPU (evaluate.py:47).  Here the clips of the evaluation set are sharded over
the ranks of a torch.distributed job (contiguous blocks, one process per GPU, full weight replica per rank),
each rank runs the fused teacher-forced evaluation (gn_teacher_forced_eval: 15 timesteps x K MaskGIT steps,
CE of the step-0 logits and accuracy of the final samples accumulated on the device), and ONE all-reduce(sum)
of 4 doubles over NCCL/NVLink produces the global numbers:

    loss = sum_CE / tokens        (== eval_utils.compute_loss averaged with AvgMetric's batch-size weights,
                                   eval_utils.py:16-25,72-77: every clip contributes the same 15*256 tokens)
    acc  = sample_correct / tokens  (evaluate.py:179)

`backend_fn` abstracts "evaluate one batch on this rank -> accumulator tensor[4]" so that the sharding and
reduction logic is testable on CPU with gloo (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced split: ranks < (n % world) get one extra item."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_metrics(acc: torch.Tensor, group=None) -> torch.Tensor:
    """all-reduce(sum) of the accumulator [sum CE, tokens, argmax-correct, sample-correct]."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def finalize(acc: torch.Tensor) -> dict:
    a = acc.detach().to("cpu", torch.float64)
    tokens = float(a[1])
    return {"loss": float(a[0]) / tokens, "acc": float(a[3]) / tokens, "argmax_acc": float(a[2]) / tokens,
            "tokens": int(tokens)}


@torch.no_grad()
def evaluate_clips(clips: torch.Tensor, backend_fn: Callable[[torch.Tensor, int], torch.Tensor], batch_size: int,
                   acc_device, rank: Optional[int] = None, world: Optional[int] = None, group=None) -> dict:
    """clips [N, T*S] (or [N,T,S]) int tokens on the host; every rank passes the SAME tensor and evaluates its
    shard.  backend_fn(batch_clips, global_index_of_first_clip) -> float64 tensor[4] on acc_device."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(clips.shape[0], rank, world)
    total = torch.zeros(4, dtype=torch.float64, device=acc_device)
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        total += backend_fn(clips[b0:b1], b0).to(acc_device)
    total = reduce_metrics(total, group)
    out = finalize(total)
    out.update(rank=rank, world=world, local_clips=hi - lo)
    return out


def b200_backend(model, maskgit_steps: int = 2, unmask_mode: str = "random", noise_seed: Optional[int] = 1234,
                 temperature: float = 0.0):
    """backend_fn for a 1xgpt_b200.STMaskGIT on this rank's GPU.  MaskGIT re-mask noise (torch.rand_like in the
    reference) is drawn per clip from a generator seeded with (noise_seed + global clip index), so results do
    not depend on how clips are sharded or batched."""
    cfg = model.config

    def fn(batch: torch.Tensor, first_index: int) -> torch.Tensor:
        B = batch.shape[0]
        noise = None
        if maskgit_steps > 1 and unmask_mode == "random":
            per_clip = []
            for i in range(B):
                g = torch.Generator().manual_seed(noise_seed + first_index + i)
                per_clip.append(torch.rand(cfg.T - 1, maskgit_steps - 1, cfg.S, generator=g))
            noise = torch.stack(per_clip, dim=2)  # [T-1, K-1, B, S]
        uniform = None
        if temperature > 1e-8:                    # Categorical draws (evaluate.py --temperature), same per-clip seeding
            per_clip = []
            for i in range(B):
                g = torch.Generator().manual_seed(noise_seed + first_index + i + (1 << 20))
                per_clip.append(torch.rand(cfg.T - 1, maskgit_steps, cfg.S, cfg.num_factored_vocabs, generator=g))
            uniform = torch.stack(per_clip, dim=2)  # [T-1, K, B, S, NV]
        return model.teacher_forced_eval(batch.reshape(B, -1), maskgit_steps=maskgit_steps, unmask_mode=unmask_mode,
                                         noise=noise, temperature=temperature, uniform=uniform)

    return fn


class GenieEvaluator:
    """Drop-in for genie/evaluate.py:68-143: `args` carries `checkpoint_dir`, `maskgit_steps`, `temperature`,
    `latent_h`, `latent_w` (the reference's argparse namespace); `decode_latents` is an instance of
    `decode_latents_wrapper()`.  `model=` hands over an already constructed 1xgpt_b200.STMaskGIT instead of loading
    `args.checkpoint_dir` (extra keyword arguments go to `STMaskGIT.from_pretrained`, e.g. precision / kv_cache).

    `predict_zframe_logits` is the reference's loop, one `maskgit_generate` per timestep (evaluate.py:103-122), for
    callers that want the samples and the factored logits; the metric loop of the CLI uses the fused
    `teacher_forced_eval` (`b200_backend` above), which never materialises the logits."""

    def __init__(self, args, decode_latents=None, device="cuda", model=None, **model_kwargs):
        from .model import STMaskGIT
        if model is None:
            model_kwargs.setdefault("kv_cache", True)
            model = STMaskGIT.from_pretrained(args.checkpoint_dir, **model_kwargs)
        self.model = model.to(device=device)
        self.model.eval()
        self.decode_latents = decode_latents
        self.device = device
        self.args = args

    @torch.no_grad()
    def predict_zframe_logits(self, input_ids: torch.LongTensor, noise: Optional[torch.Tensor] = None):
        """input_ids (B, T*H*W) -> (samples_THW (B, T-1, H, W), factored_logits (B, V, NV, T-1, H, W)).
        `noise` [T-1, K-1, B, S] optionally fixes the MaskGIT re-mask noise (torch.rand_like in the reference)."""
        cfg = self.model.config
        B = input_ids.size(0)
        window = cfg.T
        h, w = int(self.args.latent_h), int(self.args.latent_w)
        inputs_THW = input_ids.reshape(B, window, h, w).to(self.device)
        steps = int(self.args.maskgit_steps)
        all_samples, all_logits = [], []
        for timestep in range(1, window):
            inputs_masked = inputs_THW.clone()
            inputs_masked[:, timestep:] = self.model.mask_token_id
            samples_HW, factored_logits = self.model.maskgit_generate(
                inputs_masked, out_t=timestep, maskgit_steps=steps, temperature=float(self.args.temperature),
                noise=None if noise is None else noise[timestep - 1])
            all_samples.append(samples_HW)
            all_logits.append(factored_logits)
        return torch.stack(all_samples, dim=1), torch.stack(all_logits, dim=3)

    @torch.no_grad()
    def predict_next_frames(self, samples_THW) -> torch.Tensor:
        """samples (B, T-1, H, W) -> uint8 frames (B, T-1, 3, 16H, 16W) through the MAGVIT2 decoder (evaluate.py:124-143)."""
        from .eval_utils import decode_tokens
        if self.decode_latents is None:
            raise ValueError("GenieEvaluator was built without decode_latents (decode_latents_wrapper(...))")
        return decode_tokens(samples_THW.cpu(), self.decode_latents)
