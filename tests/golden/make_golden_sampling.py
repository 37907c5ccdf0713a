#!/usr/bin/env python3
"""Golden fixture for the temperature > 0 branch of maskgit_generate (st_mask_git.py:182-187), generated from the
UNMODIFIED reference (container only):   python tests/golden/make_golden_sampling.py

The reference draws `Categorical(probs=probs / temperature).sample()` from the global torch RNG, which cannot be
replayed by another implementation.  What CAN be pinned is the distribution it draws from.  For a tiny GenieConfig
with a peaked readout this script runs the reference's maskgit_generate N times per temperature (1 MaskGIT step,
fresh seed each run) and records the per-position histograms of the two factored sub-tokens, next to the step-0
logits the reference returned.  tests/test_sampling.py checks that those histograms are the ones softmax(logits)
predicts -- for EVERY temperature (Categorical renormalises probs / temperature, so the temperature cancels) --
and that the oracle's / kernel's inverse-CDF sampler reproduces the same distribution.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import genie_oracle as O  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STMaskGIT, GenieConfig, _, _ = import_reference()

TEMPS = [0.5, 1.0, 3.0]
N = 3000


@torch.no_grad()
def main():
    kw = dict(num_layers=2, num_heads=4, d_model=64, T=4, S=16, image_vocab_size=262144, num_factored_vocabs=2,
              qk_norm=False, use_mup=False)
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=51, readout_gain=300.0, bias_std=0.02)
    m = STMaskGIT(GenieConfig(**kw)).eval()
    m.load_state_dict(sd, strict=True)
    B, out_t = 2, 2
    ids = O.synthetic_clips(cfg, B, seed=151)
    prompt = ids.clone()
    prompt[:, out_t:] = cfg.mask_token_id
    V, NV, S = cfg.factored_vocab_size, cfg.num_factored_vocabs, cfg.S
    counts = np.zeros((len(TEMPS), B * S, NV, V), dtype=np.int32)
    logits0 = None
    for ti, temp in enumerate(TEMPS):
        for r in range(N):
            torch.manual_seed(100000 * ti + r)
            p = prompt.clone()
            samples, l0 = m.maskgit_generate(p, out_t, maskgit_steps=1, temperature=temp)
            logits0 = l0
            s = samples.reshape(-1).numpy()
            for i in range(NV):
                f = (s // (V ** i)) % V
                counts[ti, np.arange(B * S), i, f] += 1
        print("temperature", temp, "done")
    probs = torch.softmax(logits0, dim=1)                     # [B, V, NV, H, W]
    print("max prob per (pos, vocab): median", float(probs.amax(dim=1).median()))
    np.savez_compressed(os.path.join(OUT, "tiny_sampling.npz"), cfg=np.array(repr(kw)), seed=np.int64(51),
                        readout_gain=np.float64(300.0), ids=ids.numpy().astype(np.int32), out_t=np.int64(out_t),
                        temps=np.array(TEMPS), n_draws=np.int64(N), counts=counts, logits0=logits0.numpy())


if __name__ == "__main__":
    main()
