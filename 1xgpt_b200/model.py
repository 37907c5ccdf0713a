"""Host-side mirror of the reference's module interface for the GENIE hot path.

Same class names, constructor arguments, method signatures, state_dict keys and error behaviour as
  genie/attention.py       SelfAttention.forward(x[B',N',C], causal)           (:36, :67)
  genie/st_transformer.py  Mlp / STBlock / STTransformerDecoder.forward(tgt)    (:7-120)
  genie/factorization_utils.py FactorizedEmbedding                              (:6-52)
  genie/st_mask_git.py     STMaskGIT.{compute_logits, maskgit_generate, generate, forward,
                           from_pretrained}                                     (:29-313)
so GenieConfig JSONs and `model.safetensors` checkpoints load unchanged.  The modules below are
PARAMETER CONTAINERS: every forward goes through the C ABI of libgenie_b200.so (hand-written sm_100a
kernels); there is no PyTorch compute path and no CPU fallback.  PyTorch is used for device memory,
streams and (in evaluate.py) torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .config import GenieConfig


class ModelOutput(dict):
    """Minimal stand-in for transformers.utils.ModelOutput (attribute + key access)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def cosine_schedule(u):
    """u in [0, 1]  (genie/st_mask_git.py:17-26)"""
    if isinstance(u, torch.Tensor):
        return torch.cos(u * torch.pi / 2)
    if isinstance(u, float):
        return math.cos(u * math.pi / 2)
    raise NotImplementedError(f"Unexpected {type(u)=} {u=}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Rooted(nn.Module):
    """Sub-modules keep a weak reference to the owning STMaskGIT (which owns the native handle)."""

    def __getstate__(self):          # weakrefs are neither picklable nor deep-copyable; the owner re-links its copy
        state = dict(self.__dict__)
        state.pop("_root_ref", None)
        return state

    def _root_model(self) -> "STMaskGIT":
        ref = self.__dict__.get("_root_ref")
        root = ref() if ref is not None else None
        if root is None:
            raise RuntimeError(
                f"{type(self).__name__} is a parameter container of the B200 path; construct it through "
                "STMaskGIT (which owns the native model handle) to run it")
        return root


class SelfAttention(_Rooted):
    """genie/attention.py:9-35 (parameters) — forward runs gn_attention_forward."""

    def __init__(self, num_heads: int, d_model: int, qkv_bias: bool = False, proj_bias: bool = True,
                 qk_norm: bool = True, use_mup: bool = True, attn_drop: float = 0.0) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = d_model // num_heads
        self.scale = 8 / self.head_dim if use_mup else self.head_dim ** -0.5
        self.qkv = nn.Linear(d_model, d_model * 3, bias=qkv_bias)
        self.proj = nn.Linear(d_model, d_model, bias=proj_bias)
        self.qk_norm = qk_norm
        if self.qk_norm:
            self.norm = nn.LayerNorm(self.head_dim, eps=1e-05)

    def forward(self, x: torch.Tensor, causal: bool = False) -> torch.Tensor:
        root = self._root_model()
        layer, which = self.__dict__["_where"]
        return root._attention_forward(layer, which, x, causal)


BasicSelfAttention = SelfAttention
MemoryEfficientAttention = SelfAttention


class Mlp(nn.Module):
    """genie/st_transformer.py:7-25 (parameters only)."""

    def __init__(self, d_model: int, mlp_ratio: float = 4.0, mlp_bias: bool = True, mlp_drop: float = 0.0) -> None:
        super().__init__()
        hidden_dim = int(d_model * mlp_ratio)
        self.fc1 = nn.Linear(d_model, hidden_dim, bias=mlp_bias)
        self.fc2 = nn.Linear(hidden_dim, d_model, bias=mlp_bias)


class STBlock(nn.Module):
    """genie/st_transformer.py:28-83 (parameters only)."""

    def __init__(self, num_heads: int, d_model: int, qkv_bias: bool = False, proj_bias: bool = True,
                 qk_norm: bool = True, use_mup: bool = True, attn_drop: float = 0.0, mlp_ratio: float = 4.0,
                 mlp_bias: bool = True, mlp_drop: float = 0.0) -> None:
        super().__init__()
        self.norm1 = nn.Identity() if qk_norm else nn.LayerNorm(d_model, eps=1e-05)
        self.spatial_attn = SelfAttention(num_heads, d_model, qkv_bias, proj_bias, qk_norm, use_mup, attn_drop)
        self.temporal_attn = SelfAttention(num_heads, d_model, qkv_bias, proj_bias, qk_norm, use_mup, attn_drop)
        self.norm2 = nn.Identity() if qk_norm else nn.LayerNorm(d_model, eps=1e-05)
        self.mlp = Mlp(d_model=d_model, mlp_ratio=mlp_ratio, mlp_bias=mlp_bias, mlp_drop=mlp_drop)


class STTransformerDecoder(_Rooted):
    """genie/st_transformer.py:86-120 — forward(tgt[B,T,S,C]) -> [B,T,S,C] runs gn_decoder_forward."""

    def __init__(self, num_layers: int, num_heads: int, d_model: int, qkv_bias: bool = False,
                 proj_bias: bool = True, qk_norm: bool = True, use_mup: bool = True, attn_drop: float = 0.0,
                 mlp_ratio: float = 4.0, mlp_bias: bool = True, mlp_drop: float = 0.0):
        super().__init__()
        self.layers = nn.ModuleList([
            STBlock(num_heads, d_model, qkv_bias, proj_bias, qk_norm, use_mup, attn_drop, mlp_ratio, mlp_bias,
                    mlp_drop) for _ in range(num_layers)])

    def forward(self, tgt):
        return self._root_model()._decoder_forward(tgt)


class FactorizedEmbedding(nn.Module):
    """genie/factorization_utils.py:6-28 (parameters only; the gather is the embed kernel)."""

    def __init__(self, factored_vocab_size: int, num_factored_vocabs: int, d_model: int, mask_token_id: int):
        super().__init__()
        self.factored_vocab_size = factored_vocab_size
        self.num_factored_vocabs = num_factored_vocabs
        self.d_model = d_model
        self.mask_token_id = mask_token_id
        self.factored_embeds = nn.ParameterList([nn.Embedding(factored_vocab_size, d_model)
                                                 for _ in range(num_factored_vocabs)])
        self.mask_token_embed = nn.Parameter(torch.zeros(1, d_model))


class _NativeHandle:
    def __init__(self, cfg: _lib.gn_config, device_index: int):
        self.lib = _lib.load()
        self.ptr = C.c_void_p()
        _lib.check(self.lib.gn_model_create(C.byref(self.ptr), C.byref(cfg), device_index))

    def __del__(self):
        try:
            if self.ptr:
                self.lib.gn_model_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class STMaskGIT(nn.Module):
    """Drop-in for genie/st_mask_git.py:29 STMaskGIT (inference path).

    Extra keyword arguments (all optional, reference behaviour by default):
      precision         "fp16" (default: tcgen05 kind::f16 with IEEE fp16 operands, fp32 accumulate / residual / LN /
                        softmax; logits within 1e-3 rel of the reference at full tensor-core speed) | "bf16" (same
                        kernels, bf16 operands: fp32 range, 5e-3..8e-3 rel) | "tf32" (tcgen05 kind::tf32, fp32
                        activations) | "fp32" (CUDA-core FMA, temperature-0 ids bit-exact)
      kv_cache          temporal K/V cache + causal frame trimming for maskgit_generate / generate / evaluate:
                        bit-identical tokens, ~8-10x fewer FLOPs (reference recomputes the full window)
      chunk_tokens      tokens per L2-resident work chunk (0 = default 32768)
      generic_attention force the CUDA-core attention kernels
      fold_ln           bf16 + pre-LN configs: apply norm1 / norm2 inside the QKV / fc1 GEMM epilogues instead of a
                        separate pass (same accuracy; measured neutral on B200 because the residual GEMMs then lose
                        their 256-wide tile, so it is off by default)
      cuda_graphs       replay the per-chunk layer stack from a captured CUDA graph when running on a non-default
                        stream (default on)
      lanes             number of concurrent streams the independent clips of a MaskGIT step are dealt to (0 = library
                        default, 1 = single stream); bit-identical results for any value
    """

    def __init__(self, config: GenieConfig, precision: str = "fp16", kv_cache: bool = False, chunk_tokens: int = 0,
                 generic_attention: bool = False, fold_ln: bool = False, cuda_graphs: bool = True,
                 lanes: int = 0):
        super().__init__()
        self.h = self.w = math.isqrt(config.S)
        assert self.h ** 2 == config.S, "Expected S to be square"
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.decoder = STTransformerDecoder(
            num_layers=config.num_layers, num_heads=config.num_heads, d_model=config.d_model,
            qkv_bias=config.qkv_bias, proj_bias=config.proj_bias, qk_norm=config.qk_norm, use_mup=config.use_mup,
            attn_drop=config.attn_drop, mlp_ratio=config.mlp_ratio, mlp_bias=config.mlp_bias,
            mlp_drop=config.mlp_drop)
        self.pos_embed_TSC = torch.nn.Parameter(torch.zeros(1, config.T, config.S, config.d_model))
        self.mask_token_id = config.image_vocab_size
        self.token_embed = FactorizedEmbedding(
            factored_vocab_size=config.factored_vocab_size, num_factored_vocabs=config.num_factored_vocabs,
            d_model=config.d_model, mask_token_id=self.mask_token_id)
        self.out_x_proj = nn.Linear(config.d_model, config.factored_vocab_size * config.num_factored_vocabs)
        self.config = config
        self.precision = precision
        self.kv_cache = bool(kv_cache)
        self.chunk_tokens = int(chunk_tokens)
        self.generic_attention = bool(generic_attention)
        self.fold_ln = bool(fold_ln)
        self.cuda_graphs = bool(cuda_graphs)
        self.lanes = int(lanes)
        self.__dict__["_native"] = None
        self.__dict__["_native_key"] = None
        self.__dict__["_weights_dirty"] = True
        self._relink()
        self.requires_grad_(False)

    # ------------------------------------------------------------------ native handle management
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.__dict__["_weights_dirty"] = True
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.__dict__["_weights_dirty"] = True
        return out

    # copies own no native handle: a deep-copied / unpickled model builds its own on first use and re-links its
    # sub-modules to itself (weakrefs are atomic for copy.deepcopy and would keep pointing at the original)
    def _relink(self):
        ref = weakref.ref(self)
        self.decoder.__dict__["_root_ref"] = ref
        for i, blk in enumerate(self.decoder.layers):
            for which, att in enumerate((blk.spatial_attn, blk.temporal_attn)):
                att.__dict__["_root_ref"] = ref
                att.__dict__["_where"] = (i, which)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_native"], state["_native_key"], state["_weights_dirty"] = None, None, True
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._relink()

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._relink()
        return new

    def mark_weights_dirty(self):
        """Call after modifying parameters in place; the next forward re-uploads them."""
        self.__dict__["_weights_dirty"] = True

    @property
    def device(self):
        return self.pos_embed_TSC.device

    def _gn_config(self) -> _lib.gn_config:
        c = self.config
        return _lib.gn_config(
            num_layers=c.num_layers, num_heads=c.num_heads, d_model=c.d_model, T=c.T, S=c.S,
            image_vocab_size=c.image_vocab_size, num_factored_vocabs=c.num_factored_vocabs,
            factored_vocab_size=c.factored_vocab_size, use_mup=int(c.use_mup), qkv_bias=int(c.qkv_bias),
            proj_bias=int(c.proj_bias), qk_norm=int(c.qk_norm), mlp_bias=int(c.mlp_bias),
            mlp_ratio=float(c.mlp_ratio), precision=_lib.PRECISIONS[self.precision],
            chunk_tokens=self.chunk_tokens, kv_cache=int(self.kv_cache),
            generic_attention=int(self.generic_attention), fold_ln=int(self.fold_ln),
            cuda_graphs=int(self.cuda_graphs), lanes=self.lanes)

    def _handle(self) -> _NativeHandle:
        dev = self.device
        if dev.type != "cuda":
            raise _lib.GnError(
                "the GENIE B200 path runs on a CUDA device only (model is on "
                f"{dev}); move it with .to('cuda').  There is no CPU fallback.")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (idx, self.precision, self.kv_cache, self.chunk_tokens, self.generic_attention, self.fold_ln, self.cuda_graphs,
               self.lanes)
        d = self.__dict__
        if d["_native"] is None or d["_native_key"] != key:
            d["_native"] = _NativeHandle(self._gn_config(), idx)
            d["_native_key"] = key
            d["_weights_dirty"] = True
        if d["_weights_dirty"]:
            self._upload_weights(d["_native"], idx)
            d["_weights_dirty"] = False
        return d["_native"]

    def _upload_weights(self, h: _NativeHandle, idx: int):
        st = _stream(self.device)
        for key, t in self.state_dict().items():
            src = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
            shape = (C.c_int64 * src.dim())(*src.shape)
            _lib.check(h.lib.gn_model_set_weight(h.ptr, key.encode(), _ptr(src), shape, src.dim(), st))
        torch.cuda.current_stream(self.device).synchronize()  # sources may be temporaries
        _lib.check(h.lib.gn_model_check_weights(h.ptr))

    def _ids32(self, t: torch.Tensor, labels: bool = False) -> torch.Tensor:
        """int64 / uint32 ids of the reference -> int32 on the model's device, range-checked on the ORIGINAL dtype:
        the reference raises an IndexError for ids outside the embedding tables (factorization_utils.py:39-52) or
        the logit rows (labels); the kernels index with id % V and id / V and must never see such an id."""
        if t.numel():
            lo, hi = int(t.min()), int(t.max())
            top = self.config.image_vocab_size - (1 if labels else 0)      # inputs may hold the mask id
            if lo < 0 or hi > top:
                raise IndexError(f"token id out of range: min {lo}, max {hi}, expected [0, {top}]")
        return t.to(device=self.device, dtype=torch.int32).contiguous()

    # ------------------------------------------------------------------ module-level seams
    def _decoder_forward(self, tgt: torch.Tensor) -> torch.Tensor:
        h = self._handle()
        B, T, S, Cc = tgt.shape
        c = self.config
        if (T, S, Cc) != (c.T, c.S, c.d_model):
            raise ValueError(f"expected tgt [B,{c.T},{c.S},{c.d_model}], got {tuple(tgt.shape)}")
        x = tgt.to(device=self.device, dtype=torch.float32).contiguous()
        y = torch.empty_like(x)
        _lib.check(h.lib.gn_decoder_forward(h.ptr, _ptr(x), _ptr(y), B, _stream(self.device)))
        return y.to(tgt.dtype)

    def _attention_forward(self, layer: int, which: int, x: torch.Tensor, causal: bool) -> torch.Tensor:
        h = self._handle()
        Bq, Nq, Cc = x.shape
        xin = x.to(device=self.device, dtype=torch.float32).contiguous()
        y = torch.empty_like(xin)
        _lib.check(h.lib.gn_attention_forward(h.ptr, layer, which, _ptr(xin), _ptr(y), Bq, Nq, int(bool(causal)),
                                              _stream(self.device)))
        return y.to(x.dtype)

    # ------------------------------------------------------------------ STMaskGIT API
    def compute_logits(self, x_THW: torch.Tensor) -> torch.Tensor:
        """[B,T,H,W] ids -> [B, NV*V, T, H, W] fp32  (st_mask_git.py:255-265)"""
        h = self._handle()
        c = self.config
        B, T = x_THW.shape[0], x_THW.shape[1]
        if T != c.T or x_THW[0, 0].numel() != c.S:
            raise ValueError(f"expected ids [B,{c.T},{self.h},{self.w}], got {tuple(x_THW.shape)}")
        ids = self._ids32(x_THW.reshape(B, T, c.S))
        Cc = c.factored_vocab_size * c.num_factored_vocabs
        logits = torch.empty(B, Cc, T, self.h, self.w, device=self.device, dtype=torch.float32)
        _lib.check(h.lib.gn_compute_logits(h.ptr, _ptr(ids), B, _ptr(logits), _stream(self.device)))
        return logits

    @staticmethod
    def init_mask(prompt_THW):
        H, W = prompt_THW.size(2), prompt_THW.size(3)
        return torch.zeros(prompt_THW.size(0), H * W, dtype=torch.bool, device=prompt_THW.device)

    def _unmask_mode(self, unmask_mode: str) -> int:
        if unmask_mode == "greedy":
            return _lib.GN_UNMASK_GREEDY
        if unmask_mode == "random":
            return _lib.GN_UNMASK_RANDOM
        raise NotImplementedError(f"Expected `unmask_mode` to be one of ['greedy', 'random'], got {unmask_mode}")

    def _noise(self, shape, noise):
        """torch.rand_like of the reference (st_mask_git.py:206) is drawn by the caller-visible torch generator
        on the model's device and handed to the kernels; pass `noise` to make a run reproducible."""
        if noise is not None:
            n = noise.to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(n.shape) != tuple(shape):
                raise ValueError(f"noise must have shape {tuple(shape)}, got {tuple(n.shape)}")
            return n
        if shape[-3] == 0:
            return None
        return torch.rand(*shape, device=self.device, dtype=torch.float32)

    def _uniform(self, shape, temperature, uniform):
        """Categorical(probs / temperature).sample() of the reference (st_mask_git.py:184-187) draws from the global
        torch RNG; here the draw is an inverse CDF over uniforms [..., steps, B, S, NV] that come from the torch
        generator of the model's device, or from the caller (`uniform=`) for a reproducible run."""
        if temperature <= 1e-8:
            return None
        if uniform is not None:
            u = uniform.to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(u.shape) != tuple(shape):
                raise ValueError(f"uniform must have shape {tuple(shape)}, got {tuple(u.shape)}")
            return u
        return torch.rand(*shape, device=self.device, dtype=torch.float32)

    @torch.no_grad()
    def maskgit_generate(self, prompt_THW: torch.LongTensor, out_t: int, maskgit_steps: int = 1,
                         temperature: float = 0.0, unmask_mode: str = "random",
                         noise: Optional[torch.Tensor] = None, uniform: Optional[torch.Tensor] = None):
        """st_mask_git.py:123-229.  Mutates prompt_THW[:, out_t] in place; returns
        (sample_HW [B,H,W] int64, factored_logits [B, V, NV, H, W] of step 0)."""
        assert out_t, "maskgit_generate requires out_t > 0"
        mode = self._unmask_mode(unmask_mode)
        h = self._handle()
        c = self.config
        B, T, H, W = prompt_THW.shape
        ids = self._ids32(prompt_THW.reshape(B, T, c.S))
        if ids.data_ptr() == prompt_THW.data_ptr():
            ids = ids.clone()
        nz = None
        if mode == _lib.GN_UNMASK_RANDOM and maskgit_steps > 1:
            nz = self._noise((maskgit_steps - 1, B, c.S), noise)
        un = self._uniform((maskgit_steps, B, c.S, c.num_factored_vocabs), temperature, uniform)
        Cc = c.factored_vocab_size * c.num_factored_vocabs
        samples = torch.empty(B, c.S, device=self.device, dtype=torch.int32)
        logits0 = torch.empty(B, Cc, c.S, device=self.device, dtype=torch.float32)
        try:
            _lib.check(h.lib.gn_maskgit_generate(h.ptr, _ptr(ids), B, int(out_t), int(maskgit_steps),
                                                 float(temperature), mode, _ptr(nz), _ptr(un), _ptr(samples),
                                                 _ptr(logits0), _stream(self.device)))
        except _lib.GnError as e:
            if "must be masked" in str(e):
                raise AssertionError(f"when generating z{out_t}, frames {out_t} and later must be masked") from e
            raise
        samples_HW = samples.to(torch.long).reshape(B, H, W)
        prompt_THW[:, out_t] = samples_HW.to(prompt_THW.device)
        factored = logits0.reshape(B, c.num_factored_vocabs, c.factored_vocab_size, H, W).transpose(1, 2)
        return samples_HW, factored

    @torch.no_grad()
    def generate(self, input_ids: torch.LongTensor, attention_mask: torch.LongTensor, max_new_tokens: int,
                 min_new_tokens: int = None, return_logits: int = False, maskgit_steps: int = 1,
                 temperature: float = 0.0, noise: Optional[torch.Tensor] = None,
                 uniform: Optional[torch.Tensor] = None):
        """st_mask_git.py:65-113 (Llama-style signature; `attention_mask` ignored like the reference)."""
        assert min_new_tokens in (None, max_new_tokens), \
            "Expecting `min_new_tokens`, if specified, to match `max_new_tokens`."
        c = self.config
        assert max_new_tokens % c.S == 0, "Expecting `max_new_tokens` to be a multiple of `self.config.S`."
        num_new = max_new_tokens // c.S
        B = input_ids.size(0)
        t_prompt = input_ids.size(1) // c.S
        if t_prompt + num_new != c.T:
            raise ValueError(f"prompt frames ({t_prompt}) + new frames ({num_new}) must equal config.T ({c.T}): "
                             "pos_embed_TSC broadcasts over exactly T frames in the reference")
        h = self._handle()
        tokens = torch.full((B, c.T, c.S), self.mask_token_id, device=self.device, dtype=torch.int32)
        tokens[:, :t_prompt] = self._ids32(input_ids.reshape(B, t_prompt, c.S))
        nz = None
        if maskgit_steps > 1:
            nz = self._noise((num_new, maskgit_steps - 1, B, c.S), noise)
        Cc = c.factored_vocab_size * c.num_factored_vocabs
        logits0 = torch.empty(B, Cc, num_new, c.S, device=self.device, dtype=torch.float32) if return_logits else None
        un = self._uniform((num_new, maskgit_steps, B, c.S, c.num_factored_vocabs), temperature, uniform)
        _lib.check(h.lib.gn_generate(h.ptr, _ptr(tokens), B, t_prompt, int(maskgit_steps), float(temperature),
                                     _lib.GN_UNMASK_RANDOM, _ptr(nz), _ptr(un), _ptr(logits0), _stream(self.device)))
        predicted = tokens.to(torch.long).reshape(B, c.T * c.S)
        if return_logits:
            fl = logits0.reshape(B, c.num_factored_vocabs, c.factored_vocab_size, num_new, self.h, self.w)
            return predicted, fl.transpose(1, 2)
        return predicted

    @torch.no_grad()
    def forward(self, input_ids, labels):
        """st_mask_git.py:267-279 -> ModelOutput(loss, acc, logits[B, NV*V, T, H, W])."""
        h = self._handle()
        c = self.config
        B = input_ids.size(0)
        ids = self._ids32(input_ids.reshape(B, c.T, c.S))
        lab = self._ids32(labels.reshape(B, c.T, c.S), labels=True)
        Cc = c.factored_vocab_size * c.num_factored_vocabs
        logits = torch.empty(B, Cc, c.T, self.h, self.w, device=self.device, dtype=torch.float32)
        acc = torch.zeros(4, device=self.device, dtype=torch.float64)
        _lib.check(h.lib.gn_forward_loss(h.ptr, _ptr(ids), _ptr(lab), B, _ptr(logits), _ptr(acc),
                                         _stream(self.device)))
        loss = (acc[0] / acc[1]).to(torch.float32)
        accuracy = (acc[2] / acc[1]).to(torch.float32)
        return ModelOutput(loss=loss, acc=accuracy, logits=logits)

    @torch.no_grad()
    def compute_loss_and_acc(self, logits_CTHW, targets_THW, relevant_mask_THW):
        """st_mask_git.py:231-253: mean CE (summed over the factored vocabularies) and all-vocabularies-correct
        accuracy over the positions of frames 1.. where `relevant_mask_THW` ([B, T-1, H, W], frames 1..) is set.  `forward`
        fuses this with the readout; this stand-alone form runs the same CE kernel on caller-supplied logits
        [B, NV*V, T, H, W]."""
        from .eval_utils import factored_cross_entropy
        c = self.config
        Cc = c.factored_vocab_size * c.num_factored_vocabs
        B, T = targets_THW.shape[0], targets_THW.shape[1]
        if logits_CTHW.shape[0] != B or logits_CTHW.shape[1] != Cc or logits_CTHW.shape[2] != T:
            raise ValueError(f"expected logits [B,{Cc},T,H,W] matching targets [B,T,H,W], got "
                             f"{tuple(logits_CTHW.shape)} vs {tuple(targets_THW.shape)}")
        tg = self._ids32(targets_THW[:, 1:].reshape(B, -1), labels=True)
        rows = logits_CTHW.to(self.device)[:, :, 1:].permute(0, 2, 3, 4, 1).reshape(-1, Cc)
        if tuple(relevant_mask_THW.shape[:2]) != (B, T - 1):
            raise ValueError(f"relevant_mask_THW must cover frames 1..: [B,{T - 1},H,W], got {tuple(relevant_mask_THW.shape)}")
        w = relevant_mask_THW.to(self.device).reshape(-1)
        acc = factored_cross_entropy(rows, tg.reshape(-1), c.num_factored_vocabs, c.factored_vocab_size, weight=w)
        return (acc[0] / acc[1]).to(torch.float32), (acc[2] / acc[1]).to(torch.float32)

    @torch.no_grad()
    def teacher_forced_eval(self, input_ids: torch.Tensor, maskgit_steps: int = 2, unmask_mode: str = "random",
                            noise: Optional[torch.Tensor] = None, return_samples: bool = False,
                            temperature: float = 0.0, uniform: Optional[torch.Tensor] = None):
        """evaluate.py:82-122,173-179 fused: returns the accumulator tensor [sum CE, tokens, argmax-correct,
        sample-correct] (float64, device) for this batch, optionally the samples [B, T-1, H, W]."""
        h = self._handle()
        c = self.config
        B = input_ids.size(0)
        gt = self._ids32(input_ids.reshape(B, c.T, c.S), labels=True)
        mode = self._unmask_mode(unmask_mode)
        nz = None
        if mode == _lib.GN_UNMASK_RANDOM and maskgit_steps > 1:
            nz = self._noise((c.T - 1, maskgit_steps - 1, B, c.S), noise)
        acc = torch.zeros(4, device=self.device, dtype=torch.float64)
        samples = torch.empty(B, c.T - 1, c.S, device=self.device, dtype=torch.int32) if return_samples else None
        un = self._uniform((c.T - 1, maskgit_steps, B, c.S, c.num_factored_vocabs), temperature, uniform)
        _lib.check(h.lib.gn_teacher_forced_eval(h.ptr, _ptr(gt), B, int(maskgit_steps), float(temperature), mode,
                                                _ptr(nz), _ptr(un), _ptr(samples), _ptr(acc), _stream(self.device)))
        if return_samples:
            return acc, samples.to(torch.long).reshape(B, c.T - 1, self.h, self.w)
        return acc

    # ------------------------------------------------------------------ counters
    def flops_per_clip_forward(self) -> float:
        h = self._handle()
        return float(h.lib.gn_model_flops_per_clip_forward(h.ptr))

    def flops_executed(self) -> float:
        h = self._handle()
        return float(h.lib.gn_model_flops_executed(h.ptr))

    def bytes_executed(self) -> float:
        h = self._handle()
        return float(h.lib.gn_model_bytes_executed(h.ptr))

    def reset_counters(self):
        h = self._handle()
        h.lib.gn_model_reset_counters(h.ptr)

    # ------------------------------------------------------------------ init / checkpoint I/O
    def init_weights(self):
        """st_mask_git.py:281-296 (non-muP branch; muP only rescales the init std)."""
        std = 0.02
        for module in self.modules():
            if isinstance(module, nn.Linear):
                module.weight.data.normal_(mean=0.0, std=std)
                if module.bias is not None:
                    module.bias.data.zero_()
            elif isinstance(module, nn.Embedding):
                module.weight.data.normal_(mean=0.0, std=std)
        self.mark_weights_dirty()

    def save_pretrained(self, save_directory: str):
        """Layout of huggingface_hub.PyTorchModelHubMixin: config.json + model.safetensors."""
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(vars(self.config), f)
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()},
                  os.path.join(save_directory, "model.safetensors"))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, **kwargs):
        """Loads a directory written by the reference's `save_pretrained` (config.json + model.safetensors).
        A hub repo id is resolved through huggingface_hub when the directory does not exist (needs network)."""
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            from huggingface_hub import snapshot_download
            path = snapshot_download(path)
        with open(os.path.join(path, "config.json")) as f:
            raw = json.load(f)
        if isinstance(raw.get("config"), dict):
            raw = raw["config"]
        fields = GenieConfig.__dataclass_fields__
        config = GenieConfig(**{k: v for k, v in raw.items() if k in fields})
        model = cls(config, **kwargs)
        from safetensors.torch import load_file
        sd = load_file(os.path.join(path, "model.safetensors"))
        model.load_state_dict(sd, strict=True)
        return model
