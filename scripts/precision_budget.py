#!/usr/bin/env python3
"""Error-budget table for the operand formats of the tensor-core path (CPU, torch; no GPU needed).

Emulates the rounding points of the CUDA path on the CPU oracle: every GEMM operand (activation and weight) and
every activation that the kernels store between launches is rounded to the format under test (bf16 / fp16 / tf32),
all accumulation, the residual stream, LayerNorm, softmax statistics and GELU stay fp32 -- exactly the contract of
csrc/ (DESIGN.md section 2).  Prints the Frobenius-relative error of the logits against the fp32 oracle (= the
reference's evaluation dtype) for the three production fixtures' configurations, and, per operand role, the error
when ONLY that role is rounded (the budget table VERDICT r01 asks for).

    python scripts/precision_budget.py [--clips 1] [--configs genie35m genie138m genie138m_qknorm_mup]
"""
import argparse
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import genie_oracle as O  # noqa: E402

ROLES = ["ln_out", "raw_x", "qkv", "probs", "attn_out", "hidden", "readout_in", "w_qkv_s", "w_proj_s", "w_qkv_t",
         "w_proj_t", "w_fc1", "w_fc2", "w_readout"]


def rounder(fmt):
    if fmt == "fp32":
        return lambda t: t
    if fmt == "bf16":
        return lambda t: t.to(torch.bfloat16).to(torch.float32)
    if fmt == "fp16":
        return lambda t: t.to(torch.float16).to(torch.float32)
    if fmt == "tf32":
        def r(t):   # round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32
            i = t.contiguous().view(torch.int32)
            i = (i + 0x1000) & ~0x1FFF
            return i.view(torch.float32)
        return r
    raise ValueError(fmt)


class Q:
    """rounding per operand role; roles not in `active` stay fp32"""

    def __init__(self, fmt, active=None):
        self.r = rounder(fmt)
        self.active = set(ROLES if active is None else active)

    def __call__(self, role, t):
        return self.r(t) if role in self.active else t


def attention(sd, cfg, prefix, x_q, causal, q: Q):
    Bq, Nq, C = x_q.shape
    h, hd = cfg.num_heads, C // cfg.num_heads
    tag = "_t" if causal else "_s"
    w = q("w_qkv" + tag, sd[prefix + "qkv.weight"])
    qkv = F.linear(x_q, w, sd.get(prefix + "qkv.bias"))
    qkv = qkv.reshape(Bq, Nq, 3, h, hd).permute(2, 0, 3, 1, 4)
    qq, k, v = qkv[0], qkv[1], qkv[2]
    if cfg.qk_norm:   # applied in fp32 by the QKV epilogue, before the 16-bit rounding
        nw, nb = sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]
        qq = F.layer_norm(qq, (hd,), nw, nb, 1e-5)
        k = F.layer_norm(k, (hd,), nw, nb, 1e-5)
    qq, k, v = q("qkv", qq), q("qkv", k), q("qkv", v)
    scale = 8.0 / hd if cfg.use_mup else hd ** -0.5
    att = (qq @ k.transpose(-2, -1)) * scale
    if causal:
        keep = torch.tril(torch.ones(Nq, Nq, dtype=torch.bool))
        att = att.masked_fill(~keep, -torch.finfo(att.dtype).max)
    mx = att.amax(dim=-1, keepdim=True)
    p = torch.exp(att - mx)
    l = p.sum(dim=-1, keepdim=True)          # row sum of the un-rounded probabilities (attention_tc.cu)
    y = (q("probs", p) @ v) / l
    y = q("attn_out", y.transpose(1, 2).reshape(Bq, Nq, C))
    return F.linear(y, q("w_proj" + tag, sd[prefix + "proj.weight"]), sd.get(prefix + "proj.bias"))


def forward(sd, cfg, ids, q: Q):
    B, T = ids.shape[:2]
    x = O.embed_tokens(sd, cfg, ids.reshape(B, T, -1)) + sd["pos_embed_TSC"][:, :T]
    S, C = cfg.S, cfg.d_model
    for l in range(cfg.num_layers):
        p = f"decoder.layers.{l}."

        def norm(name, t):
            if cfg.qk_norm:
                return q("raw_x", t)
            return q("ln_out", F.layer_norm(t, (C,), sd[p + name + ".weight"], sd[p + name + ".bias"], 1e-5))

        xs = x.reshape(B * T, S, C)
        xs = xs + attention(sd, cfg, p + "spatial_attn.", norm("norm1", xs), False, q)
        xt = xs.reshape(B, T, S, C).permute(0, 2, 1, 3).reshape(B * S, T, C)
        xt = xt + attention(sd, cfg, p + "temporal_attn.", q("raw_x", xt), True, q)
        hcur = F.linear(norm("norm2", xt), q("w_fc1", sd[p + "mlp.fc1.weight"]), sd.get(p + "mlp.fc1.bias"))
        hcur = q("hidden", F.gelu(hcur))
        xt = xt + F.linear(hcur, q("w_fc2", sd[p + "mlp.fc2.weight"]), sd.get(p + "mlp.fc2.bias"))
        x = xt.reshape(B, S, T, C).permute(0, 2, 1, 3)
    xr = x * cfg.readout_input_mult if cfg.use_mup else x
    return F.linear(q("readout_in", xr), q("w_readout", sd["out_x_proj.weight"]), sd["out_x_proj.bias"])


def rel(a, b):
    return float(torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1)
    ap.add_argument("--configs", nargs="*", default=["genie35m", "genie138m", "genie138m_qknorm_mup"])
    ap.add_argument("--formats", nargs="*", default=["bf16", "fp16", "tf32"])
    ap.add_argument("--budget", action="store_true", help="also the per-role table")
    args = ap.parse_args()
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    for name in args.configs:
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
        cfg = O.OracleConfig(**ast.literal_eval(str(z["cfg"])))
        sd = O.init_state_dict(cfg, seed=int(z["seed"]), readout_gain=1.0, bias_std=0.02)
        ids = torch.from_numpy(z["ids"]).long()[:args.clips]
        with torch.no_grad():
            ref = forward(sd, cfg, ids, Q("fp32"))
            for fmt in args.formats:
                out = forward(sd, cfg, ids, Q(fmt))
                print(f"{name:24s} {fmt}: logits rel {rel(out, ref):.3e}", flush=True)
                if args.budget:
                    for role in ROLES:
                        print(f"    only {role:10s}: {rel(forward(sd, cfg, ids, Q(fmt, [role])), ref):.3e}", flush=True)
                    acts = [r for r in ROLES if not r.startswith("w_")]
                    print(f"    all activations, fp32 weights: {rel(forward(sd, cfg, ids, Q(fmt, acts)), ref):.3e}", flush=True)
                    for drop in (["w_fc1", "w_fc2"], ["w_fc1", "w_fc2", "w_readout"], ["w_qkv_s", "w_qkv_t"]):
                        keep = [r for r in ROLES if r not in drop]
                        print(f"    all but {'+'.join(drop)}: {rel(forward(sd, cfg, ids, Q(fmt, keep)), ref):.3e}", flush=True)


if __name__ == "__main__":
    main()
