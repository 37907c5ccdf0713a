"""Shared test helpers (oracle side = checker, product side = 1xgpt_b200 through the C ABI)."""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import genie_oracle as O  # noqa: E402


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return z


def golden_cfg(z):
    return ast.literal_eval(str(z["cfg"]))


def golden_sd(z):
    return {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}


def rel_fro(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


def build_b200_model(kw, sd, **opts):
    import importlib
    g = importlib.import_module("1xgpt_b200")
    m = g.STMaskGIT(g.GenieConfig(**kw), **opts)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda")
