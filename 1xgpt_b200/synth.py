"""Synthetic random-init weights with the reference's state_dict keys, for benchmarks and
smoke runs when no checkpoint is reachable: N(0, std) linears / embeddings in the style of
STMaskGIT.init_weights (genie/st_mask_git.py:281-296), optional non-zero biases and LayerNorm-affine jitter so that
every parameter takes part.  Draw order and values are identical to oracle/genie_oracle.py:init_state_dict (the tests
check that), so the CPU baseline and the B200 arm of bench.py run the same network."""
from typing import Dict

import torch

from .config import GenieConfig


def synthetic_state_dict(cfg: GenieConfig, seed: int = 0, std: float = 0.02, bias_std: float = 0.0,
                         readout_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    d, L = cfg.d_model, cfg.num_layers
    hd = d // cfg.num_heads
    hid = int(d * cfg.mlp_ratio)
    V = cfg.factored_vocab_size * cfg.num_factored_vocabs

    def n(*shape, s=std):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * s

    def b(width):
        return n(width, s=bias_std) if bias_std else torch.zeros(width)

    sd: Dict[str, torch.Tensor] = {}
    sd["pos_embed_TSC"] = n(1, cfg.T, cfg.S, d)
    sd["token_embed.mask_token_embed"] = n(1, d)
    for i in range(cfg.num_factored_vocabs):
        sd[f"token_embed.factored_embeds.{i}.weight"] = n(cfg.factored_vocab_size, d)
    for layer in range(L):
        p = f"decoder.layers.{layer}."
        for attn in ("spatial_attn", "temporal_attn"):
            sd[p + attn + ".qkv.weight"] = n(3 * d, d)
            if cfg.qkv_bias:
                sd[p + attn + ".qkv.bias"] = b(3 * d)
            sd[p + attn + ".proj.weight"] = n(d, d)
            if cfg.proj_bias:
                sd[p + attn + ".proj.bias"] = b(d)
            if cfg.qk_norm:
                sd[p + attn + ".norm.weight"] = 1.0 + b(hd)
                sd[p + attn + ".norm.bias"] = b(hd)
        if not cfg.qk_norm:
            for nm in ("norm1", "norm2"):
                sd[p + nm + ".weight"] = 1.0 + b(d)
                sd[p + nm + ".bias"] = b(d)
        sd[p + "mlp.fc1.weight"] = n(hid, d)
        sd[p + "mlp.fc2.weight"] = n(d, hid)
        if cfg.mlp_bias:
            sd[p + "mlp.fc1.bias"] = b(hid)
            sd[p + "mlp.fc2.bias"] = b(d)
    sd["out_x_proj.weight"] = n(V, d) * readout_gain
    sd["out_x_proj.bias"] = b(V)
    return sd


def synthetic_vq_state_dict(template: Dict[str, torch.Tensor], seed: int = 31) -> Dict[str, torch.Tensor]:
    """Seeded synthetic MAGVIT2 weights for throughput runs (no checkpoint reachable): fan-in scaled convolutions,
    GroupNorm affine near identity, small biases.  `template` = state_dict() of a VQModel (keys and shapes)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in template.items():
        if v.dim() == 4:
            sd[k] = torch.randn(v.shape, generator=g) / (v.shape[1] * v.shape[2] * v.shape[3]) ** 0.5
        elif "norm" in k and k.endswith(".weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
    return sd
