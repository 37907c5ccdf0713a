"""1xgpt_b200 — B200-native (sm_100a) GENIE ST-transformer + MaskGIT path behind the reference's
module interface.  `import importlib; pkg = importlib.import_module("1xgpt_b200")`, or use the
`genie_b200` alias module at the repo root.

The package needs its native library (1xgpt_b200/libgenie_b200.so, built by `make -C 1xgpt_b200/csrc`);
importing the package loads it and FAILS LOUDLY if it is missing — there is no PyTorch / CPU fallback.
"""
from . import _lib
from ._lib import GnError

_lib.load()  # fail at import time, not at first use, if the extension is absent

from .config import GenieConfig  # noqa: E402
from .model import (  # noqa: E402
    STMaskGIT, STTransformerDecoder, STBlock, Mlp, SelfAttention, BasicSelfAttention, MemoryEfficientAttention,
    FactorizedEmbedding, ModelOutput, cosine_schedule,
)
from .vq import VQModel, VQConfig, decode_latents_wrapper  # noqa: E402
from .synth import synthetic_state_dict, synthetic_vq_state_dict  # noqa: E402
from .eval_utils import AvgMetric, compute_loss, decode_tokens  # noqa: E402
from .evaluate import GenieEvaluator  # noqa: E402
from .data import RawTokenDataset, get_maskgit_collator  # noqa: E402
from .factorization_utils import factorize_token_ids, unfactorize_token_ids, factorize_labels, nth_root  # noqa: E402

__all__ = [
    "GenieConfig", "STMaskGIT", "STTransformerDecoder", "STBlock", "Mlp", "SelfAttention", "BasicSelfAttention",
    "MemoryEfficientAttention", "FactorizedEmbedding", "ModelOutput", "cosine_schedule", "factorize_token_ids",
    "unfactorize_token_ids", "factorize_labels", "nth_root", "GnError", "VQModel", "VQConfig", "decode_latents_wrapper",
    "synthetic_state_dict", "synthetic_vq_state_dict", "AvgMetric", "compute_loss", "decode_tokens", "GenieEvaluator", "RawTokenDataset",
    "get_maskgit_collator",
]
