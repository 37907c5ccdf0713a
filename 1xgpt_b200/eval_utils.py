"""The metric helpers genie/evaluate.py builds on (eval_utils.py:10-77), behind the same names and signatures.

`compute_loss` is the caller-side seam of the teacher-forced evaluation: it takes the factored logits that
`GenieEvaluator.predict_zframe_logits` returns and sums the per-vocabulary cross entropies.  Here the reduction runs in
the CE kernel of the native library (`gn_cross_entropy`, csrc/decode.cu: per-row log-sum-exp in fp32, accumulation in
float64) on the logits' CUDA device; there is no CPU implementation — logits on the host are first copied to the GPU.
(The fused evaluation, `STMaskGIT.teacher_forced_eval`, never materialises these logits at all.)

`compute_lpips` (eval_utils.py:80-87) needs the LPIPS AlexNet weights and is out of scope.
"""
from __future__ import annotations

from typing import Callable

import torch

from . import _lib


class AvgMetric:
    """Batch-size-weighted running mean (eval_utils.py:10-25)."""

    def __init__(self):
        self.total = 0
        self.count = 0

    def update(self, val, batch_size=1):
        self.total += val * batch_size
        self.count += batch_size

    def update_list(self, flat_vals):
        self.total += sum(flat_vals)
        self.count += len(flat_vals)

    def mean(self):
        return self.total / self.count if self.count else 0


def decode_tokens(reshaped_token_ids: torch.LongTensor, decode_latents: Callable) -> torch.ByteTensor:
    """(B, T, H, W) token ids -> (B, T, 3, 16H, 16W) uint8 frames (eval_utils.py:28-41).  `decode_latents` is an
    instance of `decode_latents_wrapper()`: ours returns one uint8 tensor [N,3,H,W]; a list of PIL images / arrays
    (the reference's wrapper) is accepted as well."""
    B, T = reshaped_token_ids.shape[:2]
    flat = reshaped_token_ids.reshape(B * T, *reshaped_token_ids.shape[2:]).cpu().numpy()
    decoded = decode_latents(flat)
    if not torch.is_tensor(decoded):
        import numpy as np
        decoded = torch.stack([torch.from_numpy(np.asarray(im)).permute(2, 0, 1) for im in decoded])
    return decoded.reshape(B, T, *decoded.shape[1:])


def _cuda_device(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise _lib.GnError("compute_loss runs on the GPU (gn_cross_entropy); no CUDA device is available and there is "
                           "no CPU fallback for this path")
    return torch.device("cuda", torch.cuda.current_device())


def factored_cross_entropy(logits_rows: torch.Tensor, targets: torch.Tensor, num_factored_vocabs: int,
                           factored_vocab_size: int, weight: torch.Tensor = None) -> torch.Tensor:
    """rows [R, NV*V] fp32 (vocabulary-major columns), targets [R] unfactorized ids, optional weight [R] (0 / 1)
    -> float64 tensor [sum CE, rows counted, rows whose per-vocabulary argmax all match, 0] on the rows' device."""
    dev = logits_rows.device
    R = logits_rows.shape[0]
    rows = logits_rows.to(torch.float32).contiguous()
    tg = targets.to(device=dev, dtype=torch.int32).contiguous()
    w = None if weight is None else (weight != 0).to(device=dev, dtype=torch.uint8).contiguous()
    acc = torch.zeros(4, device=dev, dtype=torch.float64)
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.gn_cross_entropy(rows.data_ptr(), tg.data_ptr(), R, int(factored_vocab_size),
                                        int(num_factored_vocabs), None if w is None else w.data_ptr(),
                                        acc.data_ptr(), st))
        torch.cuda.current_stream(dev).synchronize()      # rows / tg / w are temporaries of this call
    return acc


def compute_loss(labels_flat: torch.LongTensor, factored_logits: torch.FloatTensor, num_factored_vocabs: int = 2,
                 factored_vocab_size: int = 512) -> float:
    """Cross entropy of teacher-forced logits, summed over the factored vocabularies and averaged over tokens
    (eval_utils.py:44-77).  labels_flat (B, T*H*W); factored_logits (B, V, NV, T-1, H, W)."""
    assert factored_logits.dim() == 6 \
           and tuple(factored_logits.size()[:3]) == (labels_flat.size(0), factored_vocab_size, num_factored_vocabs), \
           f"Shape of `logits` should be (B, {factored_vocab_size}, {num_factored_vocabs}, T-1, H, W)"
    B = labels_flat.size(0)
    t = factored_logits.size(3) + 1
    h, w = factored_logits.size()[-2:]
    assert t * h * w == labels_flat.size(1), "Shape of `factored_logits` does not match flattened latent image size."
    top = factored_vocab_size ** num_factored_vocabs
    labels = labels_flat.reshape(B, t, h * w)[:, 1:]
    if labels.numel() and (int(labels.min()) < 0 or int(labels.max()) >= top):
        raise IndexError(f"label out of range [0, {top})")          # F.cross_entropy raises for these too
    dev = _cuda_device(factored_logits)
    # (B, V, NV, T-1, H, W) -> rows (B, T-1, H, W) x columns (NV, V): layout change only, the arithmetic is the kernel's
    rows = factored_logits.to(dev).permute(0, 3, 4, 5, 2, 1).reshape(-1, num_factored_vocabs * factored_vocab_size)
    acc = factored_cross_entropy(rows, labels.reshape(-1), num_factored_vocabs, factored_vocab_size)
    return float(acc[0] / acc[1])
