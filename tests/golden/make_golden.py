#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference (container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_import.py (xformers / mup stubbed, fp32, CPU,
XFORMERS_DISABLED=true) and records, for seeded synthetic weights and clips:

  tiny_*.npz        3 flag combinations of a tiny GenieConfig (full logits, MaskGIT samples with
                    injected noise, STMaskGIT.forward loss/acc, eval_utils.compute_loss,
                    STMaskGIT.generate tokens) incl. the full state_dict
  attn_*.npz        the 5 (d_model, qk_norm) cases of the reference's own test_attention.py
                    (heads=4, x=randn(1,16,d), causal=True) + non-causal, weights included
  genie35m.npz      the in-tree config genie/configs/magvit_n32_h8_d256.json, B=2 (BASELINE
                    config 1): weights are regenerated from the seed by oracle.init_state_dict
                    (sha256 stored), logits sub-sampled at fixed token positions, MaskGIT-2 samples
  genie138m.npz     GENIE_138M shape (L32 d512 h8), B=1, same content
The injected MaskGIT noise replaces torch.rand_like (st_mask_git.py:206) by monkeypatching.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import genie_oracle as O  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STMaskGIT, GenieConfig, BasicSelfAttention, ref_compute_loss = import_reference()


class inject_noise:
    """torch.rand_like -> successive rows of `noise` ([k, B, S])."""

    def __init__(self, noise):
        self.noise, self.i = noise, 0

    def __enter__(self):
        self.orig = torch.rand_like

        def fake(x, *a, **k):
            n = self.noise[self.i].reshape(x.shape).clone()
            self.i += 1
            return n

        torch.rand_like = fake
        return self

    def __exit__(self, *a):
        torch.rand_like = self.orig


def sd_sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def build_ref(kw, sd):
    m = STMaskGIT(GenieConfig(**kw)).eval()
    m.load_state_dict(sd, strict=True)
    return m


@torch.no_grad()
def tiny(name, kw, seed):
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=seed, readout_gain=4.0, bias_std=0.02)
    m = build_ref(kw, sd)
    B, K, out_t = 2, 3, 2
    ids = O.synthetic_clips(cfg, B, seed=seed + 100)
    prompt = ids.clone()
    prompt[:, out_t:] = cfg.mask_token_id
    logits = m.compute_logits(prompt)
    noise = O.tie_free_noise(K, B, cfg.S, seed=seed + 200)
    p = prompt.clone()
    with inject_noise(noise):
        samples, logits0 = m.maskgit_generate(p, out_t, maskgit_steps=K, temperature=0.0)
    # greedy unmask mode (confidence driven)
    p2 = prompt.clone()
    samples_g, _ = m.maskgit_generate(p2, out_t, maskgit_steps=K, temperature=0.0, unmask_mode="greedy")
    # forward(): MLM-style input with some masked tokens in frames >= 1
    g = torch.Generator().manual_seed(seed + 300)
    x_in = ids.clone().reshape(B, -1)
    mask = torch.rand(x_in.shape, generator=g) < 0.4
    mask[:, : cfg.S] = False
    x_in[mask] = cfg.mask_token_id
    out = m(x_in, ids.reshape(B, -1))
    # generate(): 2 prompt frames + 2 new frames, K=2
    gnoise = torch.stack([O.tie_free_noise(2, B, cfg.S, seed=seed + 400 + i) for i in range(2)])
    with inject_noise(gnoise.reshape(-1, B, cfg.S)):
        gen = m.generate(ids[:, :2].reshape(B, -1), None, max_new_tokens=2 * cfg.S, maskgit_steps=2, temperature=0.0)
    rec = {f"sd/{k}": v.numpy() for k, v in sd.items()}
    rec.update(ids=ids.numpy(), prompt=prompt.numpy(), logits=logits.numpy(), noise=noise.numpy(),
               samples=samples.numpy(), prompt_after=p.numpy(), logits0=logits0.numpy(),
               samples_greedy=samples_g.numpy(),
               fwd_in=x_in.numpy(), fwd_loss=np.float32(out.loss), fwd_acc=np.float32(out.acc),
               gen_noise=gnoise.numpy(), gen_tokens=gen.numpy(),
               cfg=np.array(repr(kw)))
    if ref_compute_loss is not None:
        # evaluate.py style: factored logits for T-1 frames from a single forward (as an extra pin of
        # eval_utils.compute_loss): [B,V,NV,T-1,H,W]
        fl = logits[:, :, 1:].reshape(B, cfg.num_factored_vocabs, cfg.factored_vocab_size, cfg.T - 1, cfg.hw, cfg.hw)
        fl = fl.transpose(1, 2).contiguous()
        rec["eval_loss"] = np.float64(ref_compute_loss(ids.reshape(B, -1), fl, cfg.num_factored_vocabs,
                                                       cfg.factored_vocab_size))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "ok", float(out.loss), float(out.acc))


@torch.no_grad()
def attn_cases():
    for d_model, qk_norm in [(32, False), (64, True), (64, False), (128, True), (128, False)]:
        torch.manual_seed(1000 + d_model + int(qk_norm))
        net = BasicSelfAttention(num_heads=4, d_model=d_model, qk_norm=qk_norm).eval()
        if qk_norm:
            net.norm.weight.data.normal_(1.0, 0.1)
            net.norm.bias.data.normal_(0.0, 0.1)
        net.proj.bias.data.normal_(0.0, 0.1)
        x = torch.randn(1, 16, d_model)
        rec = {f"sd/{k}": v.numpy() for k, v in net.state_dict().items()}
        rec.update(x=x.numpy(), y_causal=net(x, causal=True).numpy(), y_full=net(x, causal=False).numpy(),
                   use_mup=np.bool_(True))
        np.savez_compressed(os.path.join(OUT, f"attn_d{d_model}_qk{int(qk_norm)}.npz"), **rec)
    print("attn ok")


SUB_T = [1, 8, 15]
SUB_S = [0, 17, 100, 255]


@torch.no_grad()
def production(name, kw, B, seed):
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=seed, readout_gain=1.0, bias_std=0.02)
    m = build_ref(kw, sd)
    ids = O.synthetic_clips(cfg, B, seed=seed + 100)
    out_t, K = 8, 2
    prompt = ids.clone()
    prompt[:, out_t:] = cfg.mask_token_id
    # teacher-forced full window (no masks): logits for all frames
    logits_full = m.compute_logits(ids)
    noise = O.tie_free_noise(K, B, cfg.S, seed=seed + 200)
    p = prompt.clone()
    with inject_noise(noise):
        samples, logits0 = m.maskgit_generate(p, out_t, maskgit_steps=K, temperature=0.0)
    h = cfg.hw
    lf = logits_full.reshape(B, -1, cfg.T, cfg.S)
    sub = lf[:, :, SUB_T][:, :, :, SUB_S]                     # [B, C, 3, 4]
    l0 = logits0.reshape(B, cfg.factored_vocab_size, cfg.num_factored_vocabs, cfg.S)
    # top-2 margins of the step-0 logits (for the argmax-exactness bound in the GPU test)
    srt = torch.sort(l0, dim=1, descending=True).values
    margin = (srt[:, 0] - srt[:, 1])                           # [B, NV, S]
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cfg=np.array(repr(kw)), seed=np.int64(seed), sd_sha=np.array(sd_sha(sd)),
        ids=ids.numpy().astype(np.int32), noise=noise.numpy(),
        sub_t=np.array(SUB_T), sub_s=np.array(SUB_S), logits_sub=sub.numpy(),
        logits_full_absmax=np.float32(logits_full.abs().max()),
        logits_full_fro=np.float64(torch.linalg.vector_norm(logits_full.double())),
        logits0_sub=l0[:, :, :, SUB_S].numpy(), margin0=margin.numpy(),
        samples=samples.numpy().astype(np.int32), prompt_after=p.numpy().astype(np.int32),
        argmax0=(l0.argmax(dim=1)).numpy().astype(np.int32),
    )
    print(name, "ok")


if __name__ == "__main__":
    base = dict(num_layers=2, num_heads=4, d_model=64, T=4, S=16, image_vocab_size=262144, num_factored_vocabs=2)
    tiny("tiny_preln", dict(base, qk_norm=False, use_mup=False), 11)
    tiny("tiny_qknorm_mup", dict(base, qk_norm=True, use_mup=True, qkv_bias=True), 12)
    tiny("tiny_qknorm", dict(base, qk_norm=True, use_mup=False), 13)
    attn_cases()
    import json
    with open("/root/reference/genie/configs/magvit_n32_h8_d256.json") as f:
        kw35 = json.load(f)
    production("genie35m", kw35, B=2, seed=21)
    production("genie138m", dict(kw35, d_model=512), B=1, seed=22)
    production("genie138m_qknorm_mup", dict(kw35, d_model=512, qk_norm=True, use_mup=True), B=1, seed=23)


@torch.no_grad()
def magvit_fixture():
    """MAGVIT2 Encoder / LFQ / Decoder of the reference on synthetic weights and images (these
    modules import stand-alone; `lightning` is only needed by VQModel)."""
    from oracle import magvit_oracle as MO
    sys.path.insert(0, "/root/reference")
    from magvit2.config import VQConfig
    from magvit2.modules.diffusionmodules.improved_model import Encoder, Decoder
    from magvit2.modules.vqvae.lookup_free_quantize import LFQ
    cfg = MO.VQOracleConfig()
    sd = MO.init_vq_state_dict(cfg, seed=31)
    vq = VQConfig()
    enc, dec, lfq = Encoder(vq).eval(), Decoder(vq).eval(), LFQ(vq).eval()
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    g = torch.Generator().manual_seed(7)
    img = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
    z = enc(img)                                                   # [2,18,16,16]
    (quant, _, idx), _ = lfq(z, return_loss_breakdown=True)
    idx = idx.reshape(2, 16, 16)
    rec = dec(quant)                                               # VQModel.decode(quant)
    # dataset-style decode (visualize.py:111-116): little-endian tokens
    tok = torch.randint(0, 2 ** 18, (1, 16, 16), generator=g)
    q_le = lfq.get_codebook_entry(tok.reshape(1, -1), bhwc=(1, 16, 16, 18)).flip(1)
    img_le = dec(q_le)
    np.savez_compressed(
        os.path.join(OUT, "magvit.npz"), seed=np.int64(31), sd_sha=np.array(sd_sha(sd)), img_seed=np.int64(7),
        z=z.numpy(), ids=idx.numpy().astype(np.int32), quant_sign=(quant > 0).numpy(),
        rec_sub=rec[:, :, ::8, ::8].numpy(), rec_absmax=np.float32(rec.abs().max()),
        rec_fro=np.float64(torch.linalg.vector_norm(rec.double())),
        tok_le=tok.numpy().astype(np.int32), img_le_sub=img_le[:, :, ::8, ::8].numpy(),
        img_le_fro=np.float64(torch.linalg.vector_norm(img_le.double())))
    print("magvit ok", float(z.abs().mean()), float(rec.abs().max()))


if __name__ == "__main__" and os.environ.get("GOLDEN_MAGVIT", "1") == "1":
    magvit_fixture()
