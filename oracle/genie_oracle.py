"""CPU oracle for the GENIE ST-transformer + MaskGIT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`1xgpt_b200/`) imports this
file; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may.  It is a functional (state-dict in, tensors out) restatement of the
reference's PyTorch modules, written from the algorithm, in fp32 by default (the
reference's evaluation dtype) and optionally fp64 (used as "ground truth" when measuring
how far both the fp32 reference and the bf16/tf32 CUDA path are from exact arithmetic).

Parity status: the reference holds NO golden vectors for this path apart
from `test_attention.py` (a relative Basic-vs-xformers check).  This oracle is therefore
pinned against *outputs of the reference itself*, generated in the build container by
`tests/golden/make_golden.py` (which imports /root/reference with two stubs) and committed
under `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them.

Reference lines each function follows (all under /root/reference):
  attention              genie/attention.py:36-61        (BasicSelfAttention.forward)
  mlp                    genie/st_transformer.py:22-25
  st_block               genie/st_transformer.py:70-83
  decoder_forward        genie/st_transformer.py:115-120
  embed_tokens           genie/factorization_utils.py:29-52
  factorize/unfactorize  genie/factorization_utils.py:55-100
  compute_logits         genie/st_mask_git.py:255-265   (+ FixedMuReadout :316-323)
  maskgit_generate       genie/st_mask_git.py:123-229
  generate               genie/st_mask_git.py:65-113
  forward_loss_acc       genie/st_mask_git.py:231-253,267-279
  eval_compute_loss      eval_utils.py:44-77
  predict_zframe_logits  genie/evaluate.py:82-122
  cosine_schedule_n      genie/st_mask_git.py:17-26,199
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, asdict, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# config  (genie/config.py:7-55)
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    num_layers: int
    num_heads: int
    d_model: int
    T: int = 16
    S: int = 256
    image_vocab_size: int = 262144
    use_mup: bool = False
    num_factored_vocabs: int = 1
    factored_vocab_size: Optional[int] = None
    max_corrupt_rate: float = 0.2
    non_mlm_ratio: float = 0.5
    num_prompt_frames: int = 8
    qkv_bias: bool = False
    proj_bias: bool = True
    attn_drop: float = 0.0
    qk_norm: bool = True
    mlp_ratio: float = 4.0
    mlp_drop: float = 0.0
    mlp_bias: bool = True

    def __post_init__(self):
        # genie/config.py:54-55 + factorization_utils.py:103-106
        root = round(self.image_vocab_size ** (1.0 / self.num_factored_vocabs))
        if root ** self.num_factored_vocabs != self.image_vocab_size:
            raise ValueError("image_vocab_size is not a perfect power of num_factored_vocabs")
        self.factored_vocab_size = root

    @classmethod
    def from_json(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    def to_dict(self):
        return asdict(self)

    @property
    def mask_token_id(self):
        return self.image_vocab_size

    @property
    def head_dim(self):
        return self.d_model // self.num_heads

    @property
    def attn_scale(self):
        # genie/attention.py:26
        return 8.0 / self.head_dim if self.use_mup else self.head_dim ** -0.5

    @property
    def hw(self):
        h = math.isqrt(self.S)
        assert h * h == self.S
        return h

    @property
    def readout_input_mult(self):
        # FixedMuReadout (st_mask_git.py:316-323): x * output_mult / width_mult, with
        # output_mult = 1 and width_mult = d_model / 256 (base shapes hard-coded at :298-304).
        return 256.0 / self.d_model if self.use_mup else 1.0


# --------------------------------------------------------------------------------------
# synthetic weights with the reference's state_dict key names
# --------------------------------------------------------------------------------------
def init_state_dict(cfg: OracleConfig, seed: int = 0, readout_gain: float = 1.0,
                    std: float = 0.02, bias_std: float = 0.0) -> Dict[str, torch.Tensor]:
    """N(0, std) linears / embeddings (st_mask_git.py:281-296 style), optional non-zero biases
    and LN affine jitter so that every parameter participates in parity checks.
    `readout_gain` scales out_x_proj.weight to make logits peaked (documented in DESIGN.md)."""
    g = torch.Generator().manual_seed(seed)
    d, hd, L = cfg.d_model, cfg.head_dim, cfg.num_layers
    hid = int(d * cfg.mlp_ratio)
    V = cfg.factored_vocab_size * cfg.num_factored_vocabs

    def n(*shape, s=std):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * s

    sd: Dict[str, torch.Tensor] = {}
    sd["pos_embed_TSC"] = n(1, cfg.T, cfg.S, d)
    sd["token_embed.mask_token_embed"] = n(1, d)
    for i in range(cfg.num_factored_vocabs):
        sd[f"token_embed.factored_embeds.{i}.weight"] = n(cfg.factored_vocab_size, d)
    for l in range(L):
        p = f"decoder.layers.{l}."
        for attn in ("spatial_attn", "temporal_attn"):
            sd[p + attn + ".qkv.weight"] = n(3 * d, d)
            if cfg.qkv_bias:
                sd[p + attn + ".qkv.bias"] = n(3 * d, s=bias_std) if bias_std else torch.zeros(3 * d)
            sd[p + attn + ".proj.weight"] = n(d, d)
            if cfg.proj_bias:
                sd[p + attn + ".proj.bias"] = n(d, s=bias_std) if bias_std else torch.zeros(d)
            if cfg.qk_norm:
                sd[p + attn + ".norm.weight"] = 1.0 + (n(hd, s=bias_std) if bias_std else torch.zeros(hd))
                sd[p + attn + ".norm.bias"] = n(hd, s=bias_std) if bias_std else torch.zeros(hd)
        if not cfg.qk_norm:
            for nm in ("norm1", "norm2"):
                sd[p + nm + ".weight"] = 1.0 + (n(d, s=bias_std) if bias_std else torch.zeros(d))
                sd[p + nm + ".bias"] = n(d, s=bias_std) if bias_std else torch.zeros(d)
        sd[p + "mlp.fc1.weight"] = n(hid, d)
        sd[p + "mlp.fc2.weight"] = n(d, hid)
        if cfg.mlp_bias:
            sd[p + "mlp.fc1.bias"] = n(hid, s=bias_std) if bias_std else torch.zeros(hid)
            sd[p + "mlp.fc2.bias"] = n(d, s=bias_std) if bias_std else torch.zeros(d)
    sd["out_x_proj.weight"] = n(V, d) * readout_gain
    sd["out_x_proj.bias"] = n(V, s=bias_std) if bias_std else torch.zeros(V)
    return sd


def synthetic_clips(cfg: OracleConfig, batch: int, seed: int = 1234) -> torch.Tensor:
    """ids ~ U{0..image_vocab_size-1}, int64 [B,T,H,W]."""
    g = torch.Generator().manual_seed(seed)
    h = cfg.hw
    return torch.randint(0, cfg.image_vocab_size, (batch, cfg.T, h, h), generator=g, dtype=torch.int64)


def tie_free_noise(steps: int, batch: int, S: int, seed: int = 99) -> torch.Tensor:
    """[steps-1, B, S] fp32 'confidences': a random permutation / S per row, so argsort has no
    ties.  Stands in for torch.rand_like at st_mask_git.py:206."""
    g = torch.Generator().manual_seed(seed)
    k = max(steps - 1, 0)
    out = torch.empty(k, batch, S, dtype=torch.float32)
    for i in range(k):
        for b in range(batch):
            out[i, b] = torch.randperm(S, generator=g).to(torch.float32) / S
    return out


# --------------------------------------------------------------------------------------
# integer id arithmetic (factorization_utils.py:55-100)
# --------------------------------------------------------------------------------------
def factorize_token_ids(ids: torch.Tensor, nv: int, v: int) -> torch.Tensor:
    outs = []
    for i in range(nv):
        outs.append(torch.div(ids, v ** i, rounding_mode="floor") % v)
    return torch.stack(outs, dim=-1)


def unfactorize_token_ids(f: torch.Tensor, nv: int, v: int) -> torch.Tensor:
    out = torch.zeros_like(f[..., 0])
    for i in range(nv):
        out = out + f[..., i] * (v ** i)
    return out


def factorize_labels(labels_THW: torch.Tensor, nv: int, v: int) -> torch.Tensor:
    return factorize_token_ids(labels_THW, nv, v).permute(0, 4, 1, 2, 3).contiguous()


def cosine_schedule_n(step: int, steps: int, S: int) -> int:
    """tokens of frame out_t to (re-)mask after `step` (st_mask_git.py:199)."""
    return math.ceil(math.cos((step + 1) / steps * math.pi / 2) * S)


# --------------------------------------------------------------------------------------
# float path
# --------------------------------------------------------------------------------------
def _w(sd, key, dtype):
    t = sd.get(key)
    return None if t is None else t.to(dtype)


def embed_tokens(sd, cfg: OracleConfig, ids_BTS: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    v, nv, d = cfg.factored_vocab_size, cfg.num_factored_vocabs, cfg.d_model
    is_mask = ids_BTS == cfg.mask_token_id
    safe = torch.where(is_mask, torch.zeros_like(ids_BTS), ids_BTS)
    fac = factorize_token_ids(safe, nv, v)
    acc = torch.zeros(*ids_BTS.shape, d, dtype=dtype)
    # reference sums via torch.stack(...).sum(0): for nv == 2 this is e0 + e1 exactly.
    for i in range(nv):
        acc = acc + _w(sd, f"token_embed.factored_embeds.{i}.weight", dtype)[fac[..., i]]
    mask_vec = _w(sd, "token_embed.mask_token_embed", dtype).reshape(d)
    return torch.where(is_mask[..., None], mask_vec.expand_as(acc), acc)


def attention(sd, cfg: OracleConfig, prefix: str, x: torch.Tensor, causal: bool) -> torch.Tensor:
    """x [B', N', C] -> [B', N', C]; matches BasicSelfAttention (attention.py:36-61)."""
    dtype = x.dtype
    Bq, Nq, C = x.shape
    h, hd = cfg.num_heads, C // cfg.num_heads
    qkv = F.linear(x, _w(sd, prefix + "qkv.weight", dtype), _w(sd, prefix + "qkv.bias", dtype))
    qkv = qkv.reshape(Bq, Nq, 3, h, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if cfg.qk_norm:
        nw, nb = _w(sd, prefix + "norm.weight", dtype), _w(sd, prefix + "norm.bias", dtype)
        q = F.layer_norm(q, (hd,), nw, nb, 1e-5)
        k = F.layer_norm(k, (hd,), nw, nb, 1e-5)
    scale = 8.0 / hd if cfg.use_mup else hd ** -0.5
    q = q * scale
    att = q @ k.transpose(-2, -1)
    if causal:
        keep = torch.tril(torch.ones(Nq, Nq, dtype=torch.bool))
        att = att.masked_fill(~keep, -torch.finfo(att.dtype).max)
    att = att.softmax(dim=-1)
    y = (att @ v).transpose(1, 2).reshape(Bq, Nq, C)
    return F.linear(y, _w(sd, prefix + "proj.weight", dtype), _w(sd, prefix + "proj.bias", dtype))


def mlp(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    dtype = x.dtype
    hcur = F.linear(x, _w(sd, prefix + "fc1.weight", dtype), _w(sd, prefix + "fc1.bias", dtype))
    hcur = F.gelu(hcur)  # exact erf GELU (nn.GELU() default)
    return F.linear(hcur, _w(sd, prefix + "fc2.weight", dtype), _w(sd, prefix + "fc2.bias", dtype))


def st_block(sd, cfg: OracleConfig, layer: int, x_BTSC: torch.Tensor) -> torch.Tensor:
    dtype = x_BTSC.dtype
    B, T, S, C = x_BTSC.shape
    p = f"decoder.layers.{layer}."

    def norm(name, t):
        if cfg.qk_norm:
            return t  # nn.Identity (st_transformer.py:44,67)
        return F.layer_norm(t, (C,), _w(sd, p + name + ".weight", dtype), _w(sd, p + name + ".bias", dtype), 1e-5)

    xs = x_BTSC.reshape(B * T, S, C)
    xs = xs + attention(sd, cfg, p + "spatial_attn.", norm("norm1", xs), causal=False)
    xt = xs.reshape(B, T, S, C).permute(0, 2, 1, 3).reshape(B * S, T, C)
    xt = xt + attention(sd, cfg, p + "temporal_attn.", xt, causal=True)   # no LN before temporal
    xt = xt + mlp(sd, p + "mlp.", norm("norm2", xt))
    return xt.reshape(B, S, T, C).permute(0, 2, 1, 3).contiguous()


def decoder_forward(sd, cfg: OracleConfig, x_BTSC: torch.Tensor) -> torch.Tensor:
    for l in range(cfg.num_layers):
        x_BTSC = st_block(sd, cfg, l, x_BTSC)
    return x_BTSC


def hidden_states(sd, cfg: OracleConfig, ids_BTHW: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    B, T = ids_BTHW.shape[:2]
    x = embed_tokens(sd, cfg, ids_BTHW.reshape(B, T, -1), dtype)
    x = x + _w(sd, "pos_embed_TSC", dtype)[:, :T]
    return decoder_forward(sd, cfg, x)


def readout(sd, cfg: OracleConfig, x: torch.Tensor) -> torch.Tensor:
    dtype = x.dtype
    if cfg.use_mup:
        x = x * cfg.readout_input_mult
    return F.linear(x, _w(sd, "out_x_proj.weight", dtype), _w(sd, "out_x_proj.bias", dtype))


def compute_logits(sd, cfg: OracleConfig, ids_BTHW: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """-> [B, NV*V, T, H, W]  (st_mask_git.py:255-265)."""
    B, T, H, W = ids_BTHW.shape
    y = readout(sd, cfg, hidden_states(sd, cfg, ids_BTHW, dtype))         # [B,T,S,NV*V]
    return y.reshape(B, T, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


def _factored(logits_CHW: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """[B, NV*V, ...] -> [B, V, NV, ...]"""
    B = logits_CHW.shape[0]
    rest = logits_CHW.shape[2:]
    return logits_CHW.reshape(B, cfg.num_factored_vocabs, cfg.factored_vocab_size, *rest).transpose(1, 2)


def stable_argsort(x: torch.Tensor) -> torch.Tensor:
    """ascending, ties -> lower index first (the tie-break we define for st_mask_git.py:213)."""
    return torch.sort(x, dim=1, stable=True).indices


def categorical_icdf(probs_BVHW: torch.Tensor, u_BHW: torch.Tensor) -> torch.Tensor:
    """One draw per position from Categorical(probs / temperature) (st_mask_git.py:184-187).  Categorical divides
    its `probs` argument by their sum, so the temperature cancels and the distribution is `probs` itself; the draw
    is the inverse CDF at the given uniform: min{c : cumsum(p)[c] > u * sum(p)} (last index if rounding leaves none).
    The reference draws with torch.multinomial from the global RNG, which no other implementation can replay, so
    parity for this branch is (a) exact for a given uniform tensor, (b) distributional against Categorical."""
    B, V = probs_BVHW.shape[0], probs_BVHW.shape[1]
    p = probs_BVHW.reshape(B, V, -1).to(torch.float64)
    cdf = torch.cumsum(p, dim=1)
    target = u_BHW.reshape(B, 1, -1).to(torch.float64) * cdf[:, -1:, :]
    hit = cdf > target
    first = torch.where(hit.any(dim=1), hit.to(torch.int8).argmax(dim=1), torch.full_like(hit[:, 0], V - 1, dtype=torch.int64))
    return first.reshape(u_BHW.shape)


def maskgit_generate(sd, cfg: OracleConfig, prompt_THW: torch.Tensor, out_t: int, maskgit_steps: int = 1,
                     temperature: float = 0.0, unmask_mode: str = "random",
                     noise: Optional[torch.Tensor] = None, dtype=torch.float32,
                     uniform: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mutates prompt_THW[:, out_t] in place, returns (samples [B,H,W] int64, step-0 factored logits
    [B,V,NV,H,W]).  `noise` [steps-1,B,S] replaces torch.rand_like for unmask_mode='random';
    `uniform` [steps,B,S,NV] drives the Categorical draws when temperature > 1e-8 (see categorical_icdf)."""
    assert out_t, "maskgit_generate requires out_t > 0"
    if not bool(torch.all(prompt_THW[:, out_t:] == cfg.mask_token_id)):
        raise AssertionError(f"when generating z{out_t}, frames {out_t} and later must be masked")
    if temperature > 1e-8 and uniform is None:
        uniform = torch.rand(maskgit_steps, prompt_THW.shape[0], cfg.S, cfg.num_factored_vocabs)
    B, T, H, W = prompt_THW.shape
    S, V, NV = cfg.S, cfg.factored_vocab_size, cfg.num_factored_vocabs
    unmasked = torch.zeros(B, S, dtype=torch.bool)
    logits0 = None
    samples_HW = None
    for step in range(maskgit_steps):
        logits_CHW = compute_logits(sd, cfg, prompt_THW, dtype)[:, :, out_t]
        if step == 0:
            logits0 = logits_CHW.clone()
        fl = _factored(logits_CHW, cfg)                       # [B,V,NV,H,W]
        probs = torch.softmax(fl, dim=1)
        samples = torch.zeros(B, H, W, dtype=torch.int64)
        conf = torch.ones(B, H, W, dtype=probs.dtype)
        for i in reversed(range(NV)):                         # high vocab first (flip(2))
            p = probs[:, :, i]
            if temperature <= 1e-8:
                s = p.argmax(dim=1)
            else:
                s = categorical_icdf(p, uniform[step, :, :, i].reshape(B, H, W))
            samples = samples * V + s
            conf = conf * torch.gather(p, 1, s.unsqueeze(1)).squeeze(1)
        prev_unmasked = unmasked.clone()
        prev_flat = prompt_THW[:, out_t].reshape(B, S).clone()
        samples_flat = samples.reshape(B, S)
        if step != maskgit_steps - 1:
            n = cosine_schedule_n(step, maskgit_steps, S)
            if unmask_mode == "greedy":
                c = conf.reshape(B, S).to(torch.float32).clone()
            elif unmask_mode == "random":
                c = (noise[step] if noise is not None else torch.rand(B, S)).to(torch.float32).clone()
            else:
                raise NotImplementedError(unmask_mode)
            c[unmasked] = float("inf")
            order = stable_argsort(c)
            unmasked.scatter_(1, order[:, n:], True)
            samples_flat.scatter_(1, order[:, :n], cfg.mask_token_id)
        samples_flat[prev_unmasked] = prev_flat[prev_unmasked]
        samples_HW = samples_flat.reshape(B, H, W)
        prompt_THW[:, out_t] = samples_HW
    return samples_HW, _factored(logits0, cfg).contiguous()


def generate(sd, cfg: OracleConfig, input_ids: torch.Tensor, max_new_tokens: int, maskgit_steps: int = 1,
             temperature: float = 0.0, noise: Optional[torch.Tensor] = None, return_logits=False,
             dtype=torch.float32):
    """STMaskGIT.generate (st_mask_git.py:65-113).  noise: [new_frames, steps-1, B, S]."""
    assert max_new_tokens % cfg.S == 0
    new = max_new_tokens // cfg.S
    B = input_ids.shape[0]
    h = cfg.hw
    x = input_ids.clone().reshape(B, -1, h, h)
    t0 = x.shape[1]
    x = torch.cat([x, torch.full((B, new, h, h), cfg.mask_token_id, dtype=torch.int64)], dim=1)
    all_logits = []
    for i, t in enumerate(range(t0, t0 + new)):
        s, fl = maskgit_generate(sd, cfg, x, t, maskgit_steps, temperature,
                                 noise=None if noise is None else noise[i], dtype=dtype)
        x[:, t] = s
        all_logits.append(fl)
    flat = x.reshape(B, -1)
    return (flat, torch.stack(all_logits, dim=3)) if return_logits else flat


def forward_loss_acc(sd, cfg: OracleConfig, input_ids: torch.Tensor, labels: torch.Tensor, dtype=torch.float32):
    """STMaskGIT.forward (st_mask_git.py:267-279) -> (loss, acc, logits[B,NV*V,T,H,W])."""
    B = input_ids.shape[0]
    T, h = cfg.T, cfg.hw
    x = input_ids.reshape(B, T, h, h)
    lab = labels.reshape(B, T, h, h)
    logits = compute_logits(sd, cfg, x, dtype)
    relevant = (x[:, 1:] == cfg.mask_token_id)
    fl = _factored(logits[:, :, 1:], cfg)                                  # [B,V,NV,T-1,H,W]
    ft = factorize_labels(lab[:, 1:], cfg.num_factored_vocabs, cfg.factored_vocab_size)
    loss = F.cross_entropy(fl, ft, reduction="none").sum(dim=1)
    acc = (fl.argmax(dim=1) == ft).all(dim=1)
    nmask = relevant.sum()
    return (loss * relevant).sum() / nmask, (acc * relevant).sum().float() / nmask, logits


def eval_compute_loss(labels_flat: torch.Tensor, factored_logits: torch.Tensor, nv=2, v=512) -> float:
    """eval_utils.compute_loss (eval_utils.py:44-77): mean over all B*(T-1)*H*W tokens of the
    summed per-vocab CE."""
    t = factored_logits.shape[3] + 1
    h, w = factored_logits.shape[-2:]
    lab = labels_flat.reshape(labels_flat.shape[0], t, h, w)[:, 1:]
    fl = factorize_labels(lab, nv, v)
    return float(F.cross_entropy(factored_logits, fl, reduction="none").sum(dim=1).mean())


def predict_zframe_logits(sd, cfg: OracleConfig, input_ids: torch.Tensor, maskgit_steps: int = 2,
                          temperature: float = 0.0, noise: Optional[torch.Tensor] = None,
                          dtype=torch.float32):
    """GenieEvaluator.predict_zframe_logits (evaluate.py:82-122).
    noise: [T-1, steps-1, B, S].  -> samples [B,T-1,H,W], logits [B,V,NV,T-1,H,W]."""
    B = input_ids.shape[0]
    h = cfg.hw
    x = input_ids.reshape(B, cfg.T, h, h)
    samples, logits = [], []
    for i, t in enumerate(range(1, cfg.T)):
        m = x.clone()
        m[:, t:] = cfg.mask_token_id
        s, fl = maskgit_generate(sd, cfg, m, t, maskgit_steps, temperature,
                                 noise=None if noise is None else noise[i], dtype=dtype)
        samples.append(s)
        logits.append(fl)
    return torch.stack(samples, dim=1), torch.stack(logits, dim=3)


def teacher_forced_metrics(sd, cfg: OracleConfig, input_ids: torch.Tensor, maskgit_steps=2,
                           noise=None, dtype=torch.float32):
    """evaluate.py:173-179: (CE loss, accuracy, samples)."""
    samples, fl = predict_zframe_logits(sd, cfg, input_ids, maskgit_steps, 0.0, noise, dtype)
    loss = eval_compute_loss(input_ids, fl.float(), cfg.num_factored_vocabs, cfg.factored_vocab_size)
    B = input_ids.shape[0]
    gt = input_ids.reshape(B, cfg.T, cfg.hw, cfg.hw)[:, 1:]
    acc = float((gt == samples).float().mean())
    return loss, acc, samples


# FLOP model (SURVEY.md §8d / BASELINE.md §2) -------------------------------------------
def flops_per_token_layer(cfg: OracleConfig) -> int:
    d = cfg.d_model
    return 32 * d * d + 4 * d * (cfg.S + cfg.T)


def flops_per_clip_forward(cfg: OracleConfig) -> int:
    d = cfg.d_model
    v = cfg.factored_vocab_size * cfg.num_factored_vocabs
    return (cfg.num_layers * flops_per_token_layer(cfg) + 2 * d * v) * cfg.T * cfg.S
