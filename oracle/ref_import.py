"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY, and container-only: /root/reference does not exist on the GPU box,
so nothing that runs under `-m gpu`, `smoke()` or `bench.py` may call this.  It is used by
`tests/golden/make_golden.py` (to generate the committed fixtures from the reference's own
modules) and by the CPU-side test that cross-checks the oracle restatement live when the
reference happens to be present.

Two third-party imports of the reference are absent from the image and are stubbed
(SURVEY.md §8c):  `xformers.ops` (only the names; XFORMERS_DISABLED=true selects the in-tree
pure-PyTorch BasicSelfAttention, which the reference's own test_attention.py pins to the
xformers kernel at atol 1e-6) and `mup` (MuReadout as an nn.Linear subclass with
width_mult() = d_model/256 and output_mult = 1, i.e. what set_base_shapes would give for the
hard-coded base model of st_mask_git.py:298-304).
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "genie"))


def import_reference():
    """-> (STMaskGIT, GenieConfig, BasicSelfAttention, eval_compute_loss) from the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    import torch.nn as nn

    os.environ["XFORMERS_DISABLED"] = "true"
    if "xformers" not in sys.modules:
        xf = types.ModuleType("xformers")
        ops = types.ModuleType("xformers.ops")

        class LowerTriangularMask:  # noqa: D401 - name only
            pass

        def _unavailable(*a, **k):
            raise RuntimeError("xformers stub: XFORMERS_DISABLED path only")

        ops.LowerTriangularMask = LowerTriangularMask
        ops.memory_efficient_attention = _unavailable
        ops.unbind = _unavailable
        xf.ops = ops
        sys.modules["xformers"] = xf
        sys.modules["xformers.ops"] = ops
    if "mup" not in sys.modules:
        mup = types.ModuleType("mup")

        class MuReadout(nn.Linear):
            def __init__(self, *a, output_mult=1.0, **k):
                super().__init__(*a, **k)
                self.output_mult = output_mult

            def width_mult(self):
                return self.in_features / 256.0

        def set_base_shapes(model, base, rescale_params=False, **k):
            return model

        def normal_(t, mean=0.0, std=1.0):
            return t.data.normal_(mean=mean, std=std)

        mup.MuReadout = MuReadout
        mup.set_base_shapes = set_base_shapes
        mup.normal_ = normal_
        sys.modules["mup"] = mup
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from genie.st_mask_git import STMaskGIT
    from genie.config import GenieConfig
    from genie.attention import BasicSelfAttention
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_eval_utils", os.path.join(REFERENCE_ROOT, "eval_utils.py"))
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        compute_loss = mod.compute_loss
    except Exception:  # torchvision missing etc.
        compute_loss = None
    return STMaskGIT, GenieConfig, BasicSelfAttention, compute_loss
