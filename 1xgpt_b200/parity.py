"""Self-check of the CUDA path against the committed golden vectors (tests/golden/*.npz, produced by the unmodified
reference in the build container, tests/golden/make_golden*.py).  No oracle import: the fixtures carry the reference's
outputs and the seed of the synthetic weights, which `synthetic_state_dict` regenerates bit-identically (sha256 of the
state dict is stored in the fixture and checked).  Used by bench.py (the `parity` object printed next to every speed
number), scripts/parity_report.py and the GPU tests.

All metrics are against the reference's fp32 outputs:
  logits_rel               Frobenius-relative error of the sub-sampled full-window logits (compute_logits)
  logits0_max_abs          max |step-0 logits of maskgit_generate - reference| on the sub-sampled positions
  argmax_mismatches        step-0 per-vocab argmax != reference
  solid_argmax_mismatches  ... counted only where the reference's top-2 margin exceeds 4 x logits0_max_abs
                           (north star: temperature-0 ids bit-exact wherever the margin is above the numerical error)
  token_agreement          fraction of final MaskGIT-2 tokens equal to the reference's
"""
from __future__ import annotations

import ast
import hashlib
import os
from typing import Optional

import numpy as np
import torch

from .config import GenieConfig
from .model import STMaskGIT
from .synth import synthetic_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def load_fixture(name: str):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    kw = ast.literal_eval(str(z["cfg"]))
    cfg = GenieConfig(**kw)
    sd = synthetic_state_dict(cfg, seed=int(z["seed"]), bias_std=0.02)
    if _sha(sd) != str(z["sd_sha"]):
        raise RuntimeError(f"fixture {name}: regenerated weights do not match the stored sha256")
    return z, cfg, sd


def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


@torch.no_grad()
def production_parity(name: str = "genie138m", precision: str = "fp16", kv_cache: bool = True,
                      device: Optional[str] = None, model: Optional[STMaskGIT] = None) -> dict:
    """Runs the `production()` fixture `name` (make_golden.py) through the CUDA path."""
    z, cfg, sd = load_fixture(name)
    dev = torch.device(device or "cuda")
    if model is None:
        model = STMaskGIT(cfg, precision=precision, kv_cache=kv_cache)
        model.load_state_dict(sd)
        model = model.to(dev)
    ids = torch.from_numpy(z["ids"]).long()
    B = ids.shape[0]
    logits = model.compute_logits(ids.to(dev)).reshape(B, -1, cfg.T, cfg.S)
    sub = logits[:, :, z["sub_t"].tolist()][:, :, :, z["sub_s"].tolist()].cpu()
    out = {"mode": precision, "fixture": name, "kv_cache": bool(kv_cache),
           "logits_rel": rel_fro(sub, torch.from_numpy(z["logits_sub"]))}
    fro = float(torch.linalg.vector_norm(logits.double()))
    out["logits_fro_rel"] = abs(fro - float(z["logits_full_fro"])) / float(z["logits_full_fro"])
    prompt = ids.clone()
    prompt[:, 8:] = cfg.image_vocab_size
    p = prompt.to(dev)
    samples, fl = model.maskgit_generate(p, 8, maskgit_steps=2, temperature=0.0, noise=torch.from_numpy(z["noise"]))
    fl = fl.reshape(B, cfg.factored_vocab_size, cfg.num_factored_vocabs, cfg.S).cpu()
    max_abs = float((fl[:, :, :, z["sub_s"].tolist()] - torch.from_numpy(z["logits0_sub"])).abs().max())
    arg = fl.argmax(dim=1)
    ref_arg = torch.from_numpy(z["argmax0"]).long()
    margin = torch.from_numpy(z["margin0"])
    out["logits0_max_abs"] = max_abs
    out["argmax_mismatches"] = int((arg != ref_arg).sum())
    out["solid_argmax_mismatches"] = int(((arg != ref_arg) & (margin > 4 * max_abs)).sum())
    out["argmax_positions"] = int(arg.numel())
    ref_samples = torch.from_numpy(z["samples"]).long().reshape(B, -1)
    out["token_agreement"] = float((samples.reshape(B, -1).cpu() == ref_samples).float().mean())
    out["prompt_after_equal"] = bool(torch.equal(p.cpu().reshape(B, -1),
                                                 torch.from_numpy(z["prompt_after"]).long().reshape(B, -1)))
    return out


@torch.no_grad()
def eval_parity(precision: str = "fp16", kv_cache: bool = True, device: Optional[str] = None) -> dict:
    """genie138m_eval.npz: evaluate.py's teacher-forced loop on 2 clips (make_golden_r2.py)."""
    z, cfg, sd = load_fixture("genie138m_eval")
    dev = torch.device(device or "cuda")
    model = STMaskGIT(cfg, precision=precision, kv_cache=kv_cache)
    model.load_state_dict(sd)
    model = model.to(dev)
    ids = torch.from_numpy(z["ids"]).long()
    B = ids.shape[0]
    acc, samples = model.teacher_forced_eval(ids.reshape(B, -1).to(dev), maskgit_steps=2,
                                             noise=torch.from_numpy(z["noise"]), return_samples=True)
    a = acc.cpu()
    ref_samples = torch.from_numpy(z["samples"]).long()
    return {"mode": precision, "fixture": "genie138m_eval", "ce": float(a[0] / a[1]), "ce_ref": float(z["loss"]),
            "ce_abs_diff": abs(float(a[0] / a[1]) - float(z["loss"])), "acc": float(a[3] / a[1]),
            "acc_ref": float(z["acc"]), "argmax_acc": float(a[2] / a[1]), "argmax_acc_ref": float(z["argmax_acc"]),
            "token_agreement": float((samples.cpu() == ref_samples).float().mean()), "tokens": int(a[1])}


@torch.no_grad()
def gen8_parity(precision: str = "fp16", kv_cache: bool = True, device: Optional[str] = None) -> dict:
    """genie138m_gen8.npz: generate.py's 8 -> 8 frame loop on 1 clip (make_golden_r2.py)."""
    z, cfg, sd = load_fixture("genie138m_gen8")
    dev = torch.device(device or "cuda")
    model = STMaskGIT(cfg, precision=precision, kv_cache=kv_cache)
    model.load_state_dict(sd)
    model = model.to(dev)
    ids = torch.from_numpy(z["ids"]).long()
    B = ids.shape[0]
    gen, lg = model.generate(ids[:, :8].reshape(B, -1).to(dev), None, max_new_tokens=8 * cfg.S, maskgit_steps=2,
                             temperature=0.0, noise=torch.from_numpy(z["noise"]), return_logits=True)
    ref = torch.from_numpy(z["tokens"]).long()
    gen = gen.cpu()
    per_frame = (gen.reshape(B, cfg.T, cfg.S) == ref.reshape(B, cfg.T, cfg.S)).float().mean(dim=(0, 2))
    l0 = lg.reshape(B, cfg.factored_vocab_size, cfg.num_factored_vocabs, 8, cfg.S).cpu()
    # only the first generated frame sees exactly the reference's inputs (later frames depend on earlier samples)
    sub = torch.from_numpy(z["logits0_sub"])
    max_abs = float((l0[:, :, :, 0][..., [0, 17, 100, 255]] - sub[:, :, :, 0]).abs().max())
    arg = l0[:, :, :, 0].argmax(dim=1)
    ref_arg = torch.from_numpy(z["argmax0"]).long()[:, :, 0]
    margin = torch.from_numpy(z["margin0"])[:, :, 0]
    return {"mode": precision, "fixture": "genie138m_gen8", "tokens_equal": bool(torch.equal(gen, ref)),
            "token_agreement_generated": float(per_frame[8:].mean()), "token_agreement_first_frame": float(per_frame[8]),
            "first_frame_logits0_max_abs": max_abs,
            "first_frame_solid_argmax_mismatches": int(((arg != ref_arg) & (margin > 4 * max_abs)).sum())}
