#!/usr/bin/env python3
"""Golden batches of the reference's MaskGIT collator (data.py:109-169), generated from the UNMODIFIED reference
(container only):

    python tests/golden/make_golden_collator.py      ->  tests/golden/collator.npz

For each case: a small GenieConfig, seeded synthetic clips, `torch.manual_seed(s)` + `random.seed(s)`, then TWO
consecutive calls of the reference's `collate_fn` (the second call continues both RNG streams), so that a restatement must
consume the generators in exactly the reference's order.  Cases are chosen to cover the MLM branch (first masked frame
1), the non-MLM branch (random first masked frame, compounding corruption) and num_factored_vocabs 1 / 2 / 3.
"""
import importlib.util
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference, REFERENCE_ROOT  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
_, GenieConfig, _, _ = import_reference()
spec = importlib.util.spec_from_file_location("ref_data", os.path.join(REFERENCE_ROOT, "data.py"))
ref_data = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_data)

CASES = [
    # name, config kwargs, batch, seed
    ("nv2_a", dict(num_layers=1, num_heads=2, d_model=32, T=6, S=16, image_vocab_size=64, num_factored_vocabs=2,
                   num_prompt_frames=3, non_mlm_ratio=0.5, max_corrupt_rate=0.2), 3, 0),
    ("nv2_b", dict(num_layers=1, num_heads=2, d_model=32, T=6, S=16, image_vocab_size=64, num_factored_vocabs=2,
                   num_prompt_frames=3, non_mlm_ratio=0.5, max_corrupt_rate=0.2), 3, 1),
    ("nv2_mlm", dict(num_layers=1, num_heads=2, d_model=32, T=5, S=16, image_vocab_size=256, num_factored_vocabs=2,
                     num_prompt_frames=2, non_mlm_ratio=0.0, max_corrupt_rate=0.5), 2, 2),
    ("nv2_ar", dict(num_layers=1, num_heads=2, d_model=32, T=8, S=4, image_vocab_size=1024, num_factored_vocabs=2,
                    num_prompt_frames=4, non_mlm_ratio=1.0, max_corrupt_rate=0.2), 4, 3),
    ("nv1", dict(num_layers=1, num_heads=2, d_model=32, T=4, S=9, image_vocab_size=100, num_factored_vocabs=1,
                 num_prompt_frames=2, non_mlm_ratio=0.5, max_corrupt_rate=0.3), 2, 4),
    ("nv3", dict(num_layers=1, num_heads=2, d_model=32, T=4, S=16, image_vocab_size=512, num_factored_vocabs=3,
                 num_prompt_frames=1, non_mlm_ratio=0.5, max_corrupt_rate=0.2), 2, 5),
    ("prod_shape", dict(num_layers=1, num_heads=8, d_model=256, T=16, S=256, image_vocab_size=262144,
                        num_factored_vocabs=2), 2, 6),
]

rec = {"names": np.array([c[0] for c in CASES])}
branches = set()
for name, kw, B, seed in CASES:
    cfg = GenieConfig(**kw)
    g = torch.Generator().manual_seed(1000 + seed)
    clips = torch.randint(0, cfg.image_vocab_size, (B, cfg.T * cfg.S), generator=g, dtype=torch.long)
    feats = [{"input_ids": clips[i], "labels": clips[i]} for i in range(B)]
    torch.manual_seed(seed)
    random.seed(seed)
    collate = ref_data.get_maskgit_collator(cfg)
    for call in range(2):
        out = collate(feats)
        x = out["input_ids"].reshape(B, cfg.T, cfg.S)
        first = int((x == cfg.image_vocab_size).any(dim=2).any(dim=0).float().argmax())
        branches.add("mlm" if first == 1 else "ar")
        rec[f"{name}/in{call}"] = out["input_ids"].numpy()
        rec[f"{name}/lab{call}"] = out["labels"].numpy()
    rec[f"{name}/cfg"] = np.array(repr(kw))
    rec[f"{name}/clips"] = clips.numpy()
    rec[f"{name}/seed"] = np.int64(seed)
assert branches == {"mlm", "ar"}, branches
np.savez_compressed(os.path.join(OUT, "collator.npz"), **rec)
print("collator.npz ok:", len(CASES), "cases x 2 calls; branches seen:", sorted(branches))
