"""On-disk token format of the reference (data.py:17-106): `video.bin` (uint32 [num_images, s, s]), optional
`segment_ids.bin` (int32 [num_images]) and `metadata.json` (keys num_images, s, vocab_size, hz, token_dtype).
The training-side `get_maskgit_collator` (data.py:109-169, MaskGIT corruption noise for train.py) is out of scope: this
path is inference-only.  Host-side I/O only: windows of `window_size` frames taken every `stride` frames, windows that
straddle two segments dropped (`filter_interrupts`), optional de-overlapping (`filter_overlaps`)."""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset


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
