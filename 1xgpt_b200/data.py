"""On-disk token format of the reference (data.py:17-106): `video.bin` (uint32 [num_images, s, s]), optional
`segment_ids.bin` (int32 [num_images]) and `metadata.json` (keys num_images, s, vocab_size, hz, token_dtype).
Host-side I/O only: windows of `window_size` frames taken every `stride` frames, windows that
straddle two segments dropped (`filter_interrupts`), optional de-overlapping (`filter_overlaps`).

`get_maskgit_collator` (data.py:109-169) is the batch former that produces the MLM-masked / corrupted `input_ids` which
`STMaskGIT.forward(input_ids, labels)` scores (BASELINE configs[0], the in-training evaluation of train.py:262-291).  It is
host-side index work on the DataLoader's CPU tensors and consumes the torch / `random` generators in the reference's order,
so a run seeded like the reference sees the same batches (tests/test_host_logic.py pins it to reference-generated batches).
Training itself (train.py: backward, optimizer) is out of scope: this path is inference-only."""
from __future__ import annotations

import json
import math
import os
import random
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset

from .factorization_utils import factorize_token_ids, unfactorize_token_ids


class RawTokenDataset(Dataset):
    def __init__(self, data_dir, window_size, stride=1, filter_interrupts=True, filter_overlaps=False):
        data_dir = Path(data_dir)
        with open(data_dir / "metadata.json") as f:
            self.metadata = json.load(f)
        n, s = self.metadata["num_images"], self.metadata["s"]
        dtype = np.dtype(self.metadata.get("token_dtype", "uint32"))
        self.data = np.memmap(data_dir / "video.bin", dtype=dtype, mode="r", shape=(n, s, s))
        seg_path = data_dir / "segment_ids.bin"
        if os.path.isfile(seg_path):
            self.segment_ids = np.memmap(seg_path, dtype=np.int32, mode="r", shape=(n,))
        else:
            self.segment_ids = None
            if filter_interrupts:
                raise NotImplementedError("Cannot filter interrupted sequences without segment ids.")
        self.window_size, self.stride = window_size, stride
        self.video_len = (window_size - 1) * stride          # frames spanned, excluding one endpoint
        starts = np.arange(max(n - self.video_len, 0))
        if filter_interrupts and len(starts):
            seg = np.asarray(self.segment_ids)
            starts = starts[seg[starts] == seg[starts + self.video_len]]   # first and last frame in one segment
        starts = starts.tolist()
        if filter_overlaps:
            # greedy, in order: keep a start only if no already-kept start lies an exact multiple (< window) of the
            # stride before it, i.e. each frame is used by at most one window
            kept, kept_set = [], set()
            for st in starts:
                if not any((st - i * stride) in kept_set for i in range(1, window_size)):
                    kept.append(st)
                    kept_set.add(st)
            starts = kept
        self.valid_start_inds = starts

    def __len__(self):
        return len(self.valid_start_inds)

    def __getitem__(self, idx):
        st = self.valid_start_inds[idx]
        x = torch.from_numpy(np.asarray(self.data[st: st + self.video_len + 1: self.stride]).astype(np.int64)).flatten()
        return {"input_ids": x, "labels": x, "attention_mask": torch.ones_like(x)}

    def clips(self, indices=None) -> torch.Tensor:
        """[N, window_size * s * s] int64 tensor of the selected (default: all) windows."""
        idx = range(len(self)) if indices is None else indices
        return torch.stack([self[i]["input_ids"] for i in idx]) if len(idx) else torch.empty(0, dtype=torch.int64)


def get_maskgit_collator(config):
    """features (list of {"input_ids": [T*S] ids}) -> {"input_ids": corrupted + masked [B, T*S], "labels": clean [B, T*S]}.

    Per batch (data.py:113-167): (1) every factored sub-token is replaced by a random one with probability
    `max_corrupt_rate * u`, u ~ U[0,1) drawn once; (2) with probability `non_mlm_ratio` the batch imitates autoregressive
    inference: frames before a random `first` in [num_prompt_frames, T-1] stay clean and every later frame is corrupted
    further at a rate that grows frame by frame, otherwise `first = 1`; (3) each frame >= `first` of each clip is masked
    at its own cosine-schedule rate (redrawn until at least one token is masked); masked positions take the mask id."""
    T, NV, V = config.T, config.num_factored_vocabs, config.factored_vocab_size
    side = math.isqrt(config.S)
    mask_id = config.image_vocab_size

    def compound_corruption(fact, replacement, first, dev):
        # later frames keep fewer of their sub-tokens: the kept fraction shrinks by U(0.9, 1) per frame
        keep = random.uniform(0.25, 1.0)
        for t in range(first, T):
            keep *= random.uniform(0.9, 1.0)
            swap = torch.rand((fact.shape[0], side, side, NV), device=dev) > keep
            fact[:, t][swap] = replacement[:, t][swap]

    def draw_mask(like, first):
        # per (clip, frame) masking probability cos(pi/2 * u); at least one masked token per batch
        while True:
            prob = torch.cos(torch.rand(like.shape[0], T - first, 1, 1) * torch.pi / 2)
            mask = torch.rand_like(like[:, first:], dtype=torch.float) < prob
            if mask.max() != 0:
                return mask

    def collate_fn(features):
        ids = torch.stack([ex["input_ids"] for ex in features])
        dev = ids.device
        clean = ids.reshape(len(features), T, side, side)
        fact = factorize_token_ids(clean, NV, V)
        noisy = torch.rand(fact.size(), device=dev) < config.max_corrupt_rate * torch.rand((), device=dev)
        replacement = torch.randint(low=0, high=V, size=fact.size(), dtype=torch.long, device=dev)
        fact[noisy] = replacement[noisy]
        first = 1
        if random.random() < config.non_mlm_ratio:
            first = random.randint(config.num_prompt_frames, T - 1)
            compound_corruption(fact, replacement, first, dev)
        mask = draw_mask(clean, first)
        x = unfactorize_token_ids(fact, NV, V)
        x[:, first:][mask] = mask_id
        return {"input_ids": x.reshape(len(features), -1), "labels": clean.clone().reshape(len(features), -1)}

    return collate_fn
