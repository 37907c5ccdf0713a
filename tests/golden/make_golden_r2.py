#!/usr/bin/env python3
"""Round-2 golden fixtures, generated from the UNMODIFIED reference (container only):

    python tests/golden/make_golden_r2.py

  genie35m_fwd.npz    BASELINE.json configs[0]: STMaskGIT.forward (st_mask_git.py:267-279) of the in-tree 35M config
                      on 2 synthetic clips with an MLM-style masked input: loss, acc, sub-sampled logits
  genie138m_eval.npz  evaluate.py's temporally teacher-forced loop (evaluate.py:103-122) on the GENIE_138M shape,
                      2 clips, MaskGIT-2: the reference's maskgit_generate is called for t = 1..15 exactly like
                      GenieEvaluator.predict_zframe_logits does, then eval_utils.compute_loss (CE) and
                      evaluate.py:179 (acc); per-timestep CE too
  genie138m_gen8.npz  generate.py's loop (generate.py:77-103): 8 prompt frames -> 8 generated frames, MaskGIT-2,
                      temperature 0, 1 clip, through STMaskGIT.generate (st_mask_git.py:65-113)
  ckpt_tiny/          a checkpoint directory written by the REFERENCE's save_pretrained (PyTorchModelHubMixin:
                      config.json + model.safetensors) for a tiny GenieConfig, + ckpt_tiny_expected.npz: tokens the
                      reference generates from it for the synthetic dataset the CLI test builds (generate.py loop,
                      MaskGIT-1 so that no RNG is involved) and its teacher-forced CE / acc (K = 1)

The injected MaskGIT noise replaces torch.rand_like (st_mask_git.py:206) by monkeypatching, as in make_golden.py.
"""
import json
import os
import shutil
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import genie_oracle as O  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STMaskGIT, GenieConfig, _, ref_compute_loss = import_reference()
sys.path.insert(0, OUT)
from make_golden import inject_noise, sd_sha, build_ref, SUB_T, SUB_S  # noqa: E402  (helpers only; guarded mains)

with open("/root/reference/genie/configs/magvit_n32_h8_d256.json") as f:
    KW35 = json.load(f)


@torch.no_grad()
def fwd_35m():
    seed = 21                                           # same weights / clips as genie35m.npz
    cfg = O.OracleConfig(**KW35)
    sd = O.init_state_dict(cfg, seed=seed, readout_gain=1.0, bias_std=0.02)
    m = build_ref(KW35, sd)
    B = 2
    ids = O.synthetic_clips(cfg, B, seed=seed + 100)
    g = torch.Generator().manual_seed(seed + 300)
    x_in = ids.clone().reshape(B, -1)
    mask = torch.rand(x_in.shape, generator=g) < 0.4
    mask[:, : cfg.S] = False
    x_in[mask] = cfg.mask_token_id
    out = m(x_in, ids.reshape(B, -1))
    lf = out.logits.reshape(B, -1, cfg.T, cfg.S)
    np.savez_compressed(os.path.join(OUT, "genie35m_fwd.npz"), cfg=np.array(repr(KW35)), seed=np.int64(seed),
                        sd_sha=np.array(sd_sha(sd)), ids=ids.numpy().astype(np.int32),
                        fwd_in=x_in.numpy().astype(np.int32), fwd_loss=np.float64(out.loss), fwd_acc=np.float64(out.acc),
                        n_masked=np.int64(mask.sum()), sub_t=np.array(SUB_T), sub_s=np.array(SUB_S),
                        logits_sub=lf[:, :, SUB_T][:, :, :, SUB_S].numpy(),
                        logits_full_fro=np.float64(torch.linalg.vector_norm(out.logits.double())))
    print("genie35m_fwd ok", float(out.loss), float(out.acc))


@torch.no_grad()
def eval_138m():
    seed, B, K = 22, 2, 2
    kw = dict(KW35, d_model=512)
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=seed, readout_gain=1.0, bias_std=0.02)
    m = build_ref(kw, sd)
    ids = O.synthetic_clips(cfg, B, seed=seed + 500)
    noise = torch.stack([O.tie_free_noise(K, B, cfg.S, seed=seed + 600 + t) for t in range(cfg.T - 1)])  # [15,1,B,S]
    # evaluate.py:103-122
    all_samples, all_logits = [], []
    for i, t in enumerate(range(1, cfg.T)):
        inputs_masked = ids.clone()
        inputs_masked[:, t:] = m.mask_token_id
        with inject_noise(noise[i]):
            samples_HW, factored_logits = m.maskgit_generate(inputs_masked, out_t=t, maskgit_steps=K, temperature=0.0)
        all_samples.append(samples_HW)
        all_logits.append(factored_logits)
    samples = torch.stack(all_samples, dim=1)                  # [B, 15, H, W]
    fl = torch.stack(all_logits, dim=3)                        # [B, V, NV, 15, H, W]
    loss = ref_compute_loss(ids.reshape(B, -1), fl, cfg.num_factored_vocabs, cfg.factored_vocab_size)   # evaluate.py:177
    acc = float((ids[:, 1:] == samples).float().mean())                                                 # evaluate.py:179
    ft = O.factorize_labels(ids[:, 1:], cfg.num_factored_vocabs, cfg.factored_vocab_size)
    ce_t = F.cross_entropy(fl, ft, reduction="none").sum(dim=1).mean(dim=(0, 2, 3))                     # per timestep
    argmax_acc = float((fl.argmax(dim=1) == ft).all(dim=1).float().mean())
    np.savez_compressed(os.path.join(OUT, "genie138m_eval.npz"), cfg=np.array(repr(kw)), seed=np.int64(seed),
                        sd_sha=np.array(sd_sha(sd)), ids=ids.numpy().astype(np.int32), noise=noise.numpy(),
                        loss=np.float64(loss), acc=np.float64(acc), argmax_acc=np.float64(argmax_acc),
                        ce_per_timestep=ce_t.numpy(), samples=samples.numpy().astype(np.int32))
    print("genie138m_eval ok", loss, acc)


@torch.no_grad()
def gen8_138m():
    seed, B, K, t_prompt = 22, 1, 2, 8
    kw = dict(KW35, d_model=512)
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=seed, readout_gain=1.0, bias_std=0.02)
    m = build_ref(kw, sd)
    ids = O.synthetic_clips(cfg, B, seed=seed + 700)
    n_new = cfg.T - t_prompt
    noise = torch.stack([O.tie_free_noise(K, B, cfg.S, seed=seed + 800 + i) for i in range(n_new)])     # [8,1,B,S]
    with inject_noise(noise.reshape(-1, B, cfg.S)):
        gen, lg = m.generate(ids[:, :t_prompt].reshape(B, -1), None, max_new_tokens=n_new * cfg.S, maskgit_steps=K,
                             temperature=0.0, return_logits=True)
    # lg: [B, V, NV, n_new, H, W] step-0 logits of every generated frame
    l0 = lg.reshape(B, cfg.factored_vocab_size, cfg.num_factored_vocabs, n_new, cfg.S)
    srt = torch.sort(l0, dim=1, descending=True).values
    np.savez_compressed(os.path.join(OUT, "genie138m_gen8.npz"), cfg=np.array(repr(kw)), seed=np.int64(seed),
                        sd_sha=np.array(sd_sha(sd)), ids=ids.numpy().astype(np.int32), noise=noise.numpy(),
                        tokens=gen.numpy().astype(np.int32), margin0=(srt[:, 0] - srt[:, 1]).numpy(),
                        argmax0=l0.argmax(dim=1).numpy().astype(np.int32),
                        logits0_sub=l0[..., SUB_S].numpy())
    print("genie138m_gen8 ok")


def synthetic_dataset(side, n_frames=700, boundary=300, seed=5, vocab=262144):
    """The token stream the CLI test writes to video.bin (two segments), identical here and in the test
    (tests/test_gpu_cli.py holds a copy of this function)."""
    g = np.random.default_rng(seed)
    video = g.integers(0, vocab, size=(n_frames, side, side), dtype=np.uint32)
    seg = np.zeros(n_frames, dtype=np.int32)
    seg[boundary:] = 1
    return video, seg


@torch.no_grad()
def ckpt_tiny():
    kw = dict(num_layers=2, num_heads=4, d_model=64, T=16, S=16, image_vocab_size=262144, num_factored_vocabs=2,
              qk_norm=False, use_mup=False)
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=61, readout_gain=4.0, bias_std=0.02)
    m = build_ref(kw, sd)
    d = os.path.join(OUT, "ckpt_tiny")
    shutil.rmtree(d, ignore_errors=True)
    m.save_pretrained(d)                                      # the reference's own writer (PyTorchModelHubMixin)
    for fn in os.listdir(d):
        if fn not in ("config.json", "model.safetensors"):
            os.remove(os.path.join(d, fn))                    # README.md model card: not part of the format we read
    m2 = STMaskGIT.from_pretrained(d).eval()                  # and its own reader: round trip
    # the dataset directory the CLI test builds, read with the REFERENCE's RawTokenDataset (data.py:17-106):
    # generate.py:66-71 (window 16, stride 15, example 0) and evaluate.py:150 (filter_overlaps=True)
    import tempfile
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_data", "/root/reference/data.py")
    ref_data = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_data)
    video, seg = synthetic_dataset(4)
    with tempfile.TemporaryDirectory() as td:
        video.tofile(os.path.join(td, "video.bin"))
        seg.tofile(os.path.join(td, "segment_ids.bin"))
        with open(os.path.join(td, "metadata.json"), "w") as f:
            json.dump({"num_images": int(video.shape[0]), "s": 4, "vocab_size": 262144, "hz": 2,
                       "token_dtype": "uint32"}, f)
        ds_gen = ref_data.RawTokenDataset(td, window_size=16, stride=15)
        starts = list(ds_gen.valid_start_inds)
        ex = ds_gen[0]["input_ids"].reshape(1, 16, 4, 4).clone()
        ds_eval = ref_data.RawTokenDataset(td, window_size=16, stride=15, filter_overlaps=True)
        kept = list(ds_eval.valid_start_inds)
        clips = torch.stack([ds_eval[i]["input_ids"] for i in range(3)]).reshape(3, 16, 4, 4).clone()
    # generate.py:77-103 with maskgit_steps=1 (no RNG): 8 prompt frames
    prompt = ex.clone()
    prompt[:, 8:] = m2.mask_token_id
    for t in range(8, 16):
        s_hw, _ = m2.maskgit_generate(prompt, out_t=t, maskgit_steps=1, temperature=0.0)
        prompt[:, t] = s_hw
    # evaluate.py:103-122, K = 1, first three non-overlapping windows
    all_s, all_l = [], []
    for t in range(1, 16):
        im = clips.clone()
        im[:, t:] = m2.mask_token_id
        s_hw, flg = m2.maskgit_generate(im, out_t=t, maskgit_steps=1, temperature=0.0)
        all_s.append(s_hw)
        all_l.append(flg)
    samples = torch.stack(all_s, dim=1)
    fl = torch.stack(all_l, dim=3)
    loss = ref_compute_loss(clips.reshape(clips.shape[0], -1), fl, 2, 512)
    acc = float((clips[:, 1:] == samples).float().mean())
    np.savez_compressed(os.path.join(OUT, "ckpt_tiny_expected.npz"), cfg=np.array(repr(kw)), first_start=np.int64(starts[0]),
                        n_valid=np.int64(len(starts)), example=ex.numpy().astype(np.int32),
                        generated=prompt.numpy().astype(np.int32), eval_starts=np.array(kept[:3]),
                        eval_loss=np.float64(loss), eval_acc=np.float64(acc), n_kept=np.int64(len(kept)))
    print("ckpt_tiny ok", loss, acc, len(starts), len(kept))


def magvit_ckpt():
    """A Lightning-format MAGVIT2 checkpoint ({"state_dict": ...}) with the key layout the reference's VQModel produces
    (lfqgan.py:21-119): generator keys `encoder.*` / `decoder.*`, EMA shadow buffers `model_ema.<name without dots>`
    written by the REFERENCE's LitEma (ema.py:20-26) + its decay / num_updates, and loss / discriminator keys that an
    inference loader must drop.  Tiny VQConfig so the file stays ~1 MB."""
    sys.path.insert(0, "/root/reference")
    from magvit2.config import VQConfig
    from magvit2.modules.diffusionmodules.improved_model import Encoder, Decoder
    from magvit2.modules.ema import LitEma
    torch.manual_seed(71)
    vq = VQConfig(base_channels=32, ch_mult=(1, 1), num_res_blocks=1)

    class Gen(torch.nn.Module):          # the parameter-holding part of lfqgan.VQModel (LFQ has no persistent state)
        def __init__(self):
            super().__init__()
            self.encoder, self.decoder = Encoder(vq), Decoder(vq)
            self.loss = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 1, 3))   # stands for loss.*

    g = Gen()
    ema = LitEma(g)
    with torch.no_grad():
        for name, buf in ema.named_buffers():
            if buf.dtype.is_floating_point and buf.dim() > 0:
                buf.mul_(0.5).add_(0.01)          # EMA weights differ from the live ones
    sd = dict(g.state_dict())
    sd.update({"model_ema." + k: v for k, v in ema.state_dict().items()})
    torch.save({"state_dict": sd, "epoch": 3, "global_step": 100}, os.path.join(OUT, "magvit_tiny.ckpt"))
    with open(os.path.join(OUT, "magvit_tiny_keys.json"), "w") as f:
        json.dump({"keys": sorted(sd), "ema_map": ema.m_name2s_name,
                   "config": {"base_channels": 32, "ch_mult": [1, 1], "num_res_blocks": 1}}, f)
    print("magvit_tiny.ckpt ok", len(sd))


if __name__ == "__main__":
    which = sys.argv[1:] or ["fwd35", "ckpt", "gen8", "eval138", "magvit_ckpt"]
    if "magvit_ckpt" in which:
        magvit_ckpt()
    if "fwd35" in which:
        fwd_35m()
    if "ckpt" in which:
        ckpt_tiny()
    if "gen8" in which:
        gen8_138m()
    if "eval138" in which:
        eval_138m()
