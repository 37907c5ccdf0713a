"""MAGVIT2 tokenizer behind the reference's module names (parameter containers + C ABI calls).

  magvit2/config.py:12-18                  VQConfig (the architecture fields)
  magvit2/models/lfqgan.py:121-129         VQModel.encode / decode
  magvit2/modules/diffusionmodules/improved_model.py   Encoder / Decoder / ResBlock / Upsampler (state_dict keys)
  visualize.py:84-122                      decode_latents_wrapper (little-endian dataset tokens -> uint8 frames)

`VQModel.state_dict()` has exactly the reference's `encoder.*` / `decoder.*` keys, so a converted `magvit2.ckpt`
state dict loads with strict=True.  Training-only parts (entropy / commitment losses, discriminator, EMA) are not
part of this path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .model import _ptr, _stream


@dataclass
class VQConfig:
    in_channels: int = 3
    z_channels: int = 18
    out_channels: int = 3
    base_channels: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4)
    num_res_blocks: int = 2
    num_codebooks: int = 1
    codebook_size: int = 262144
    token_factorization: bool = False


class ResBlock(nn.Module):
    def __init__(self, in_filters, out_filters):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_filters, eps=1e-6)
        self.norm2 = nn.GroupNorm(32, out_filters, eps=1e-6)
        self.conv1 = nn.Conv2d(in_filters, out_filters, (3, 3), padding=1, bias=False)
        self.conv2 = nn.Conv2d(out_filters, out_filters, (3, 3), padding=1, bias=False)
        if in_filters != out_filters:
            self.nin_shortcut = nn.Conv2d(in_filters, out_filters, (1, 1), padding=0, bias=False)


class Encoder(nn.Module):
    def __init__(self, config: VQConfig):
        super().__init__()
        nb = len(config.ch_mult)
        self.conv_in = nn.Conv2d(config.in_channels, config.base_channels, (3, 3), padding=1, bias=False)
        self.down = nn.ModuleList()
        in_ch_mult = (1,) + tuple(config.ch_mult)
        block_in = config.base_channels
        for i in range(nb):
            block_in = config.base_channels * in_ch_mult[i]
            block_out = config.base_channels * config.ch_mult[i]
            down = nn.Module()
            down.block = nn.ModuleList()
            for _ in range(config.num_res_blocks):
                down.block.append(ResBlock(block_in, block_out))
                block_in = block_out
            if i < nb - 1:
                down.downsample = nn.Conv2d(block_out, block_out, (3, 3), stride=(2, 2), padding=1)
            self.down.append(down)
        self.mid_block = nn.ModuleList([ResBlock(block_in, block_in) for _ in range(config.num_res_blocks)])
        self.norm_out = nn.GroupNorm(32, block_in, eps=1e-6)
        self.conv_out = nn.Conv2d(block_in, config.z_channels, (1, 1))


class Upsampler(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv1 = nn.Conv2d(dim, dim * 4, (3, 3), padding=1)


class Decoder(nn.Module):
    def __init__(self, config: VQConfig):
        super().__init__()
        nb = len(config.ch_mult)
        block_in = config.base_channels * config.ch_mult[nb - 1]
        self.conv_in = nn.Conv2d(config.z_channels, block_in, (3, 3), padding=1, bias=True)
        self.mid_block = nn.ModuleList([ResBlock(block_in, block_in) for _ in range(config.num_res_blocks)])
        self.up = nn.ModuleList()
        for i in reversed(range(nb)):
            block_out = config.base_channels * config.ch_mult[i]
            up = nn.Module()
            up.block = nn.ModuleList()
            for _ in range(config.num_res_blocks):
                up.block.append(ResBlock(block_in, block_out))
                block_in = block_out
            if i > 0:
                up.upsample = Upsampler(block_in)
            self.up.insert(0, up)
        self.norm_out = nn.GroupNorm(32, block_in, eps=1e-6)
        self.conv_out = nn.Conv2d(block_in, config.out_channels, (3, 3), padding=1)


class _VqHandle:
    def __init__(self, cfg: "_lib.gn_vq_config", device_index: int):
        self.lib = _lib.load()
        self.ptr = C.c_void_p()
        _lib.check(self.lib.gn_vq_create(C.byref(self.ptr), C.byref(cfg), device_index))

    def __del__(self):
        try:
            if self.ptr:
                self.lib.gn_vq_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class VQModel(nn.Module):
    """Inference part of magvit2/models/lfqgan.py VQModel: encode -> tokens, tokens -> decode."""

    def __init__(self, config: Optional[VQConfig] = None, precision: str = "fp16"):
        super().__init__()
        if precision not in ("fp16", "bf16", "fp32"):
            raise ValueError(f"precision must be 'fp16' (default), 'bf16' or 'fp32' (exact mode), got {precision!r}")
        self.precision = precision
        self.config = config or VQConfig()
        self.encoder = Encoder(self.config)
        self.decoder = Decoder(self.config)
        self.codebook_dim = self.config.z_channels
        self.__dict__["_native"] = None
        self.__dict__["_native_dev"] = None
        self.__dict__["_dirty"] = True
        self.requires_grad_(False)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.__dict__["_dirty"] = True
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.__dict__["_dirty"] = True
        return out

    @property
    def device(self):
        return self.encoder.conv_in.weight.device

    def _handle(self) -> _VqHandle:
        dev = self.device
        if dev.type != "cuda":
            raise _lib.GnError(f"the MAGVIT2 B200 path runs on a CUDA device only (model is on {dev}); "
                               "there is no CPU fallback")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        d = self.__dict__
        if d["_native"] is None or d["_native_dev"] != (idx, self.precision):
            c = self.config
            mult = (C.c_int32 * 8)(*(list(c.ch_mult) + [0] * (8 - len(c.ch_mult))))
            cfg = _lib.gn_vq_config(in_channels=c.in_channels, z_channels=c.z_channels, out_channels=c.out_channels,
                                    base_channels=c.base_channels, num_blocks=len(c.ch_mult), ch_mult=mult,
                                    num_res_blocks=c.num_res_blocks, precision=_lib.PRECISIONS[self.precision])
            d["_native"] = _VqHandle(cfg, idx)
            d["_native_dev"] = (idx, self.precision)
            d["_dirty"] = True
        if d["_dirty"]:
            h = d["_native"]
            st = _stream(dev)
            for key, t in self.state_dict().items():
                src = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                shape = (C.c_int64 * src.dim())(*src.shape)
                _lib.check(h.lib.gn_vq_set_weight(h.ptr, key.encode(), _ptr(src), shape, src.dim(), st))
            torch.cuda.current_stream(dev).synchronize()
            d["_dirty"] = False
        return d["_native"]

    # ------------------------------------------------------------------ checkpoint I/O
    def init_from_ckpt(self, path, stage: Optional[str] = None):
        """magvit2/models/lfqgan.py:85-119 for inference: `path` is a Lightning checkpoint ({"state_dict": ...}) of the
        reference's VQModel.  Its state dict holds the generator (`encoder.*`, `decoder.*`), the GAN / perceptual loss
        (`loss.*`, dropped), nothing for the LFQ quantizer (all its buffers are non-persistent), and the EMA shadow
        weights as buffers `model_ema.<parameter name with the dots removed>` (ema.py:20-26).
          stage=None           the reference's `decode_latents_wrapper` path (visualize.py:100-101): plain resume.  The
                               reference constructs its LitEma AFTER loading (lfqgan.py:56-57), so the `model_ema.*`
                               keys of the file are ignored and the live `encoder.* / decoder.*` weights are used.
          stage="transformer"  the EMA generator weights are loaded instead (lfqgan.py:91-110)."""
        sd = torch.load(path, map_location="cpu", weights_only=False)
        sd = sd["state_dict"] if "state_dict" in sd else sd
        own = self.state_dict()
        new = {}
        if stage == "transformer":
            ema = {k[len("model_ema."):]: v for k, v in sd.items() if k.startswith("model_ema.")}
            for k in own:
                s_name = k.replace(".", "")
                if s_name not in ema:
                    raise KeyError(f"checkpoint has no EMA weight model_ema.{s_name} for {k}")
                new[k] = ema[s_name]
        else:
            for k in own:
                if k not in sd:
                    raise KeyError(f"checkpoint is missing {k}")
                new[k] = sd[k]
        self.load_state_dict(new, strict=True)
        return self

    @classmethod
    def from_ckpt(cls, path, config: Optional[VQConfig] = None, stage: Optional[str] = None,
                  precision: str = "fp16") -> "VQModel":
        return cls(config, precision=precision).init_from_ckpt(path, stage=stage)

    # ------------------------------------------------------------------ API
    @torch.no_grad()
    def encode_to_tokens(self, x: torch.Tensor, return_latents: bool = False):
        """x [B,3,H,W] in [-1,1] -> LFQ indices [B, H/16, W/16] (big-endian, lookup_free_quantize.py:257)."""
        h = self._handle()
        B, Cc, H, W = x.shape
        down = 2 ** (len(self.config.ch_mult) - 1)
        img = x.to(device=self.device, dtype=torch.float32).contiguous()
        ids = torch.empty(B, H // down, W // down, device=self.device, dtype=torch.int32)
        z = torch.empty(B, self.config.z_channels, H // down, W // down, device=self.device) if return_latents else None
        _lib.check(h.lib.gn_vq_encode(h.ptr, _ptr(img), B, H, W, _ptr(ids), _ptr(z), _stream(self.device)))
        ids = ids.to(torch.long)
        return (ids, z) if return_latents else ids

    @torch.no_grad()
    def encode(self, x):
        """lfqgan.py:121-125 signature: (quant, emb_loss, info, loss_breakdown); the losses are training-only."""
        ids = self.encode_to_tokens(x)
        Z = self.config.z_channels
        mask = 2 ** torch.arange(Z - 1, -1, -1, device=ids.device)
        quant = ((ids.unsqueeze(1) & mask.view(1, Z, 1, 1)) != 0).float() * 2.0 - 1.0
        return quant, torch.zeros((), device=ids.device), ids.flatten(), None

    @torch.no_grad()
    def decode_tokens(self, ids: torch.Tensor, little_endian: bool = True, as_uint8: bool = False) -> torch.Tensor:
        """ids [B,h,w] -> image [B,3,16h,16w]; little_endian=True is the dataset convention (visualize.py:115)."""
        h = self._handle()
        B, hh, ww = ids.shape
        up = 2 ** (len(self.config.ch_mult) - 1)
        t = ids.to(device=self.device, dtype=torch.int32).contiguous()
        f32 = None if as_uint8 else torch.empty(B, 3, hh * up, ww * up, device=self.device, dtype=torch.float32)
        u8 = torch.empty(B, 3, hh * up, ww * up, device=self.device, dtype=torch.uint8) if as_uint8 else None
        _lib.check(h.lib.gn_vq_decode(h.ptr, _ptr(t), B, hh, ww, int(bool(little_endian)), _ptr(f32), _ptr(u8),
                                      _stream(self.device)))
        return u8 if as_uint8 else f32

    @torch.no_grad()
    def decode(self, quant: torch.Tensor) -> torch.Tensor:
        """lfqgan.py:127-129: quant [B,Z,h,w] of +-1 -> image."""
        Z = self.config.z_channels
        mask = 2 ** torch.arange(Z - 1, -1, -1, device=quant.device)
        ids = ((quant > 0).long() * mask.view(1, Z, 1, 1)).sum(dim=1)
        return self.decode_tokens(ids, little_endian=False)


def decode_latents_wrapper(model=None, batch_size: int = 16, tokenizer_ckpt: Optional[str] = None, device="cuda"):
    """visualize.py:95-122 without PIL: video_data (b,h,w) integer tokens -> uint8 tensor [b,3,H,W] on the host.
    Either pass a VQModel, or `tokenizer_ckpt=` (the reference's signature: a Lightning `magvit2.ckpt`)."""
    if model is None or isinstance(model, (str, bytes)):
        path = tokenizer_ckpt if model is None else model
        if path is None:
            raise ValueError("decode_latents_wrapper needs a VQModel or tokenizer_ckpt=")
        model = VQModel.from_ckpt(path).to(device)

    @torch.no_grad()
    def decode_latents(video_data) -> torch.Tensor:
        data = torch.as_tensor(video_data.astype("int64") if hasattr(video_data, "astype") else video_data)
        outs = []
        for s in range(0, data.shape[0], batch_size):
            outs.append(model.decode_tokens(data[s:s + batch_size], little_endian=True, as_uint8=True).cpu())
        return torch.cat(outs)

    return decode_latents
