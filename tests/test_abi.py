"""The C-ABI library loads on a CPU-only box and exports every symbol include/genie_b200.h declares."""
import ctypes
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "genie_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    fns = header_functions()
    for must in ["gn_model_create", "gn_decoder_forward", "gn_compute_logits", "gn_maskgit_generate", "gn_generate",
                 "gn_generate_host", "gn_teacher_forced_eval", "gn_forward_loss", "gn_attention_forward",
                 "gn_linear_forward", "gn_last_error", "gn_version"]:
        assert must in fns


def test_library_exports_every_declared_symbol():
    pkg = importlib.import_module("1xgpt_b200")
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/genie_b200.h but not exported"


def test_python_binding_covers_header():
    pkg = importlib.import_module("1xgpt_b200")
    assert sorted(pkg._lib.SIGNATURES) == header_functions()
    hdr = open(os.path.join(ROOT, "include", "genie_b200.h")).read()
    assert pkg._lib.load().gn_version() == int(re.search(r"#define GN_ABI_VERSION (\d+)", hdr).group(1)) == 4
    # struct mirrors: one ctypes field per C struct member (gn_config / gn_vq_config are passed by pointer)
    for cname, cls in (("gn_config", pkg._lib.gn_config), ("gn_vq_config", pkg._lib.gn_vq_config)):
        body = re.search(r"typedef struct " + cname + r" \{(.*?)\} " + cname + ";", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S),
                         re.S).group(1)
        members = re.findall(r"\b(?:int32_t|float)\s+([A-Za-z_0-9]+)", body)
        assert members == [f[0] for f in cls._fields_], (cname, members)


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "genie_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code
    assert re.findall(r"#include\s*[<\"]([^>\"]+)", code) == ["stdint.h"]
