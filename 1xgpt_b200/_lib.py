"""ctypes binding of libgenie_b200.so (the C ABI declared in include/genie_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgenie_b200.so")

GN_PREC_BF16, GN_PREC_TF32, GN_PREC_FP32, GN_PREC_FP16 = 0, 1, 2, 3
GN_UNMASK_RANDOM, GN_UNMASK_GREEDY = 0, 1
PRECISIONS = {"bf16": GN_PREC_BF16, "tf32": GN_PREC_TF32, "fp32": GN_PREC_FP32, "fp16": GN_PREC_FP16}


class GnError(RuntimeError):
    pass


class gn_config(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("num_heads", C.c_int32), ("d_model", C.c_int32), ("T", C.c_int32),
        ("S", C.c_int32), ("image_vocab_size", C.c_int32), ("num_factored_vocabs", C.c_int32),
        ("factored_vocab_size", C.c_int32), ("use_mup", C.c_int32), ("qkv_bias", C.c_int32),
        ("proj_bias", C.c_int32), ("qk_norm", C.c_int32), ("mlp_bias", C.c_int32), ("mlp_ratio", C.c_float),
        ("precision", C.c_int32), ("chunk_tokens", C.c_int32), ("kv_cache", C.c_int32),
        ("generic_attention", C.c_int32), ("fold_ln", C.c_int32), ("cuda_graphs", C.c_int32),
        ("lanes", C.c_int32),
    ]


class gn_vq_config(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("z_channels", C.c_int32), ("out_channels", C.c_int32),
                ("base_channels", C.c_int32), ("num_blocks", C.c_int32), ("ch_mult", C.c_int32 * 8),
                ("num_res_blocks", C.c_int32), ("precision", C.c_int32)]


_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); must list every symbol of include/genie_b200.h (checked by tests/test_abi.py)
SIGNATURES = {
    "gn_version": (_i, []),
    "gn_last_error": (C.c_char_p, []),
    "gn_model_create": (_i, [C.POINTER(_vp), C.POINTER(gn_config), _i]),
    "gn_model_destroy": (None, [_vp]),
    "gn_model_set_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _vp]),
    "gn_model_check_weights": (_i, [_vp]),
    "gn_decoder_forward": (_i, [_vp, _vp, _vp, _i, _vp]),
    "gn_attention_forward": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "gn_compute_logits": (_i, [_vp, _vp, _i, _vp, _vp]),
    "gn_maskgit_generate": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "gn_generate": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "gn_generate_host": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp]),
    "gn_teacher_forced_eval": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "gn_forward_loss": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "gn_linear_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "gn_spatial_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    "gn_sample_tokens": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gn_remask_step": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "gn_cross_entropy": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "gn_vq_create": (_i, [C.POINTER(_vp), C.POINTER(gn_vq_config), _i]),
    "gn_vq_destroy": (None, [_vp]),
    "gn_vq_set_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _vp]),
    "gn_vq_check_weights": (_i, [_vp, _i, _i]),
    "gn_vq_encode": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "gn_vq_decode": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "gn_profile_begin": (_i, []),
    "gn_profile_end": (_i, [C.POINTER(C.c_double)]),
    "gn_kernel_launches": (C.c_uint64, []),
    "gn_fallback_launches": (C.c_uint64, []),
    "gn_model_flops_per_clip_forward": (C.c_double, [_vp]),
    "gn_model_flops_executed": (C.c_double, [_vp]),
    "gn_model_bytes_executed": (C.c_double, [_vp]),
    "gn_model_reset_counters": (None, [_vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises GnError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GnError(
            f"{LIB_PATH} not found: build it with `make -C 1xgpt_b200/csrc` (or `python -c 'import "
            f"__graft_entry__ as g; g.build()'`).  There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().gn_last_error()
        raise GnError(f"libgenie_b200 error {rc}: {msg.decode() if msg else ''}")
