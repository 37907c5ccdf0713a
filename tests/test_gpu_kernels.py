"""GPU parity of the individual kernels (called through the C ABI) against fp32 references."""
import ctypes as C
import importlib
import math

import numpy as np
import pytest
import torch

from helpers import O, rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    g = importlib.import_module("1xgpt_b200")
    return g._lib.load(), g._lib


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def run_linear(lib, M, N, K, epi, in_bf16, out_bf16, dual=False, bias=True, simt=False, seed=0):
    L, _l = lib
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    r = torch.randn(M, N, device="cuda", generator=g) if epi == 2 else None
    if in_bf16:
        a_in, w_in = a.bfloat16().contiguous(), w.bfloat16().contiguous()
        a_ref, w_ref = a_in.float(), w_in.float()
    else:
        a_in, w_in = a, w
        a_ref, w_ref = a, w
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    out2 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16) if dual else None
    _l.check(L.gn_linear_forward(P(a_in), P(w_in), P(b), P(r), P(out), P(out2), M, N, K, epi, int(in_bf16),
                                 int(out_bf16), int(simt), None))
    torch.cuda.synchronize()
    ref = a_ref.double() @ w_ref.double().t()
    if b is not None:
        ref = ref + b.double()
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = ref + r.double()
    return out, out2, ref


# (M, N, K): production shapes of d=512 / d=256 / d=1024 layers plus ragged M and small K
TC_SHAPES = [(256, 1536, 512), (4096, 512, 512), (1000, 2048, 512), (384, 512, 2048), (128, 768, 256),
             (300, 1024, 1024), (128, 64, 64), (130, 192, 64), (8192, 1024, 512), (256, 4096, 1024)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_linear_bf16_store(lib, M, N, K):
    out, _, ref = run_linear(lib, M, N, K, epi=0, in_bf16=True, out_bf16=True)
    assert torch.isfinite(out.float()).all()
    assert rel_fro(out.float(), ref) < 4e-3          # bf16 output rounding (2^-9 per element)
    out, _, ref = run_linear(lib, M, N, K, epi=0, in_bf16=True, out_bf16=False)
    assert rel_fro(out, ref) < 2e-5                  # exact bf16 products, fp32 accumulation


@pytest.mark.parametrize("M,N,K", TC_SHAPES[:6])
def test_linear_bf16_gelu_and_residual(lib, M, N, K):
    out, _, ref = run_linear(lib, M, N, K, epi=1, in_bf16=True, out_bf16=True)
    assert rel_fro(out.float(), ref) < 4e-3
    out, out2, ref = run_linear(lib, M, N, K, epi=2, in_bf16=True, out_bf16=False, dual=True)
    assert rel_fro(out, ref) < 2e-5
    assert torch.equal(out2, out.bfloat16())         # the bf16 copy is the rounding of the fp32 output
    out, _, ref = run_linear(lib, M, N, K, epi=2, in_bf16=True, out_bf16=False, dual=False, bias=False)
    assert rel_fro(out, ref) < 2e-5


@pytest.mark.parametrize("M,N,K", TC_SHAPES[:6])
def test_linear_tf32(lib, M, N, K):
    for epi in (0, 1, 2):
        out, _, ref = run_linear(lib, M, N, K, epi=epi, in_bf16=False, out_bf16=False)
        assert rel_fro(out, ref) < 1e-3              # tf32 operands: 2^-11 relative per product


@pytest.mark.parametrize("M,N,K", [(64, 96, 32), (17, 32, 8), (128, 128, 64), (300, 1024, 1024)])
def test_linear_simt_matches(lib, M, N, K):
    for epi in (0, 1, 2):
        out, out2, ref = run_linear(lib, M, N, K, epi=epi, in_bf16=False, out_bf16=False, simt=True, dual=(epi == 2))
        assert rel_fro(out, ref) < 1e-6
        out, _, ref = run_linear(lib, M, N, K, epi=epi, in_bf16=True, out_bf16=True, simt=True)
        assert rel_fro(out.float(), ref) < 4e-3


def test_linear_tensor_vs_simt_bitwise_close(lib):
    # same bf16 inputs through both device paths: differences only from fp32 summation order
    o1, _, _ = run_linear(lib, 512, 512, 512, epi=0, in_bf16=True, out_bf16=False, simt=False, seed=3)
    o2, _, _ = run_linear(lib, 512, 512, 512, epi=0, in_bf16=True, out_bf16=False, simt=True, seed=3)
    assert rel_fro(o1, o2) < 1e-6


# ------------------------------------------------------------------------------------- decode kernels
def test_sample_tokens_bit_exact(lib):
    L, _l = lib
    torch.manual_seed(0)
    R, V, NV = 777, 512, 2
    logits = (torch.randn(R, NV * V) * 3).cuda()
    samples = torch.empty(R, dtype=torch.int32, device="cuda")
    conf = torch.empty(R, dtype=torch.float32, device="cuda")
    _l.check(L.gn_sample_tokens(P(logits), R, V, NV, None, P(samples), P(conf), None))
    lc = logits.cpu().reshape(R, NV, V)
    probs = torch.softmax(lc, dim=2)
    ids = torch.zeros(R, dtype=torch.int64)
    cf = torch.ones(R)
    for i in reversed(range(NV)):
        s = probs[:, i].argmax(dim=1)
        ids = ids * V + s
        cf = cf * probs[:, i].gather(1, s[:, None])[:, 0]
    assert torch.equal(samples.cpu().long(), ids)                       # integer ids: bit exact
    assert torch.allclose(conf.cpu(), cf, rtol=2e-5, atol=0)


def test_sample_tokens_tie_breaks_low_index(lib):
    L, _l = lib
    logits = torch.zeros(4, 1024, device="cuda")
    logits[1, 7] = logits[1, 300] = 2.0          # tie in vocab 0 -> 7
    logits[2, 512 + 9] = logits[2, 512 + 8] = 1.0  # tie in vocab 1 -> 8
    samples = torch.empty(4, dtype=torch.int32, device="cuda")
    conf = torch.empty(4, dtype=torch.float32, device="cuda")
    _l.check(L.gn_sample_tokens(P(logits), 4, 512, 2, None, P(samples), P(conf), None))
    assert samples.cpu().tolist() == [0, 7, 8 * 512, 0]


@pytest.mark.parametrize("S,steps", [(256, 2), (256, 8), (16, 3), (64, 5)])
def test_remask_bit_exact_vs_oracle_schedule(lib, S, steps):
    """Drive the remask kernel with the oracle's per-step samples/noise and compare every intermediate."""
    L, _l = lib
    B, mask_id = 5, 262144
    g = torch.Generator().manual_seed(S + steps)
    noise = O.tie_free_noise(steps, B, S, seed=S * 7 + steps)
    frame = torch.full((B, S), mask_id, dtype=torch.int64)
    unmasked = torch.zeros(B, S, dtype=torch.bool)
    d_frame = frame.to(torch.int32).cuda()
    d_unm = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
    for step in range(steps):
        samples = torch.randint(0, mask_id, (B, S), generator=g)
        # oracle (st_mask_git.py:192-223)
        prev_unm, prev = unmasked.clone(), frame.clone()
        sf = samples.clone()
        last = step == steps - 1
        n = 0
        if not last:
            n = O.cosine_schedule_n(step, steps, S)
            c = noise[step].clone()
            c[unmasked] = float("inf")
            order = O.stable_argsort(c)
            unmasked.scatter_(1, order[:, n:], True)
            sf.scatter_(1, order[:, :n], mask_id)
        sf[prev_unm] = prev[prev_unm]
        frame = sf
        # kernel
        d_s = samples.to(torch.int32).cuda()
        d_out = torch.empty(B, S, dtype=torch.int32, device="cuda")
        nz = noise[step].cuda().contiguous() if not last else None
        _l.check(L.gn_remask_step(P(d_frame), S, P(d_s), P(nz), P(d_unm), P(d_out), B, S, n, int(last), mask_id, None))
        torch.cuda.synchronize()
        assert torch.equal(d_out.cpu().long(), frame)
        assert torch.equal(d_frame.cpu().long(), frame)
        if not last:
            assert torch.equal(d_unm.cpu().bool(), unmasked)
    assert (frame != mask_id).all()


def test_remask_stable_ties(lib):
    L, _l = lib
    B, S, mask_id = 1, 32, 99
    conf = torch.zeros(B, S, device="cuda")             # all equal -> order = index
    frame = torch.full((B, S), mask_id, dtype=torch.int32, device="cuda")
    unm = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
    samples = torch.arange(S, dtype=torch.int32, device="cuda").reshape(B, S)
    out = torch.empty_like(samples)
    _l.check(L.gn_remask_step(P(frame), S, P(samples), P(conf), P(unm), P(out), B, S, 10, 0, mask_id, None))
    exp = torch.arange(S)
    exp[:10] = mask_id
    assert out.cpu()[0].tolist() == exp.tolist()


def test_cross_entropy_matches_torch(lib):
    L, _l = lib
    torch.manual_seed(1)
    R, V, NV = 1000, 512, 2
    logits = (torch.randn(R, NV * V) * 2).cuda()
    tgt = torch.randint(0, V ** NV, (R,))
    w = (torch.rand(R) < 0.6).to(torch.uint8)
    tgt_d = tgt.to(torch.int32).cuda()          # keep device buffers alive across the (async) call
    w_d = w.cuda()
    for weight in (None, w):
        acc = torch.zeros(4, dtype=torch.float64, device="cuda")
        _l.check(L.gn_cross_entropy(P(logits), P(tgt_d), R, V, NV, P(w_d) if weight is not None else None, P(acc),
                                    None))
        lc = logits.cpu().reshape(R, NV, V).double()
        f = O.factorize_token_ids(tgt, NV, V)
        ce = sum(torch.nn.functional.cross_entropy(lc[:, i], f[:, i], reduction="none") for i in range(NV))
        ok = torch.stack([lc[:, i].argmax(1) == f[:, i] for i in range(NV)]).all(0)
        sel = torch.ones(R, dtype=torch.bool) if weight is None else weight.bool()
        a = acc.cpu()
        assert a[1].item() == sel.sum().item()
        assert abs(a[0].item() - ce[sel].sum().item()) < 1e-3 * R
        assert a[2].item() == ok[sel].sum().item()


def test_sample_tokens_categorical_inverse_cdf(lib):
    """temperature > 0 branch (st_mask_git.py:182-187): inverse-CDF draw from caller-supplied uniforms.  Ids must
    equal the float64 inverse CDF of the same logits except where u*sum lies within fp32 rounding of a CDF step."""
    L, _l = lib
    torch.manual_seed(1)
    R, V, NV = 4096, 512, 2
    logits = (torch.randn(R, NV * V) * 4).cuda()
    u = torch.rand(R, NV, generator=torch.Generator().manual_seed(2))
    u[0] = 0.0
    u[1] = 0.99999994
    samples = torch.empty(R, dtype=torch.int32, device="cuda")
    conf = torch.empty(R, dtype=torch.float32, device="cuda")
    _l.check(L.gn_sample_tokens(P(logits), R, V, NV, P(u.cuda()), P(samples), P(conf), None))
    lc = logits.cpu().reshape(R, NV, V).double()
    probs = torch.softmax(lc, dim=2)
    cdf = torch.cumsum(probs, dim=2)
    got = samples.cpu().long()
    cf = torch.ones(R, dtype=torch.float64)
    mism = 0
    for i in range(NV):
        f = (got // (V ** i)) % V
        hit = cdf[:, i] > u[:, i:i + 1].double()
        ref = torch.where(hit.any(dim=1), hit.to(torch.int8).argmax(dim=1), torch.full((R,), V - 1))
        bad = f != ref
        # a mismatch is only legitimate at a CDF step: |cdf[candidate] - u| tiny for the lower of the two candidates
        lo = torch.minimum(f, ref)
        gap = (cdf[:, i].gather(1, lo[:, None])[:, 0] - u[:, i].double()).abs()
        assert bool((gap[bad] < 1e-5).all())
        assert bool(((f - ref).abs()[bad] <= 1).all())
        mism += int(bad.sum())
        cf = cf * probs[:, i].gather(1, f[:, None])[:, 0]
    print("categorical boundary mismatches:", mism, "of", R * NV)
    assert mism <= 4
    assert torch.allclose(conf.cpu().double(), cf, rtol=2e-5, atol=0)
    # distribution: 4096 draws from ONE row's distribution follow softmax (chi-square z-score)
    row = (torch.randn(1, NV * V) * 3).repeat(R, 1).cuda()
    _l.check(L.gn_sample_tokens(P(row), R, V, NV, P(u.cuda()), P(samples), P(conf), None))
    p0 = torch.softmax(row[0].cpu().double().reshape(NV, V), dim=1)
    f0 = (samples.cpu().long() % V)
    counts = torch.bincount(f0[2:], minlength=V).double()
    exp = p0[0] * (R - 2)
    big = exp >= 5
    chi2 = float(((counts[big] - exp[big]) ** 2 / exp[big]).sum() +
                 (counts[~big].sum() - exp[~big].sum()) ** 2 / exp[~big].sum())
    dof = int(big.sum())
    assert abs(chi2 - dof) / math.sqrt(2 * dof) < 5


def _spatial_ref(qkv, n_frames, S, H, hd, scale):
    """softmax(q k^T * scale) v per (frame, head) in float64 from the bf16 inputs (attention.py:48-58, non-causal)."""
    d = H * hd
    x = qkv.double().reshape(n_frames, S, 3, H, hd).permute(2, 0, 3, 1, 4)      # [3, F, H, S, hd]
    q, k, v = x[0], x[1], x[2]
    p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(n_frames * S, d)


@pytest.mark.parametrize("variant", ["2", "1"])
@pytest.mark.parametrize("n_frames,S,H,peaked", [(3, 256, 8, False), (2, 128, 8, False), (5, 256, 16, True),
                                                  (1, 256, 2, True), (47, 256, 8, False)])
def test_spatial_attention_tcgen05(lib, monkeypatch, n_frames, S, H, peaked, variant):
    """tcgen05 spatial attention (TMEM-resident scores, P fed to the PV MMA from TMEM, MN-major V operand) against a
    float64 reference and against the mma.sync kernel it replaces.  `peaked`: large score range (exercises the
    max-subtraction); each output element is a distinct mix of V rows, so a wrong key/column mapping cannot pass.
    variant 2 = persistent warp-specialised kernel (S = 256; 47 x 8 items > 148 CTAs exercises the stage ring and
    the barrier phases over several items per CTA), 1 = one CTA per 128-query tile."""
    monkeypatch.setenv("GENIE_B200_SPATIAL_TC", variant)
    L, _l = lib
    hd = 64
    d = H * hd
    g = torch.Generator(device="cuda").manual_seed(100 + n_frames + S + H)
    qkv = torch.randn(n_frames * S, 3 * d, device="cuda", generator=g)
    if peaked:
        qkv[:, :2 * d] *= 3.0
    qkv = qkv.bfloat16().contiguous()
    scale = 1.0 / math.sqrt(hd)
    outs = []
    for kernel in (0, 1):
        out = torch.full((n_frames * S, d), float("nan"), device="cuda", dtype=torch.bfloat16)
        _l.check(L.gn_spatial_attention(P(qkv), P(out), n_frames, S, H, hd, scale, kernel, None))
        torch.cuda.synchronize()
        outs.append(out)
    ref = _spatial_ref(qkv, n_frames, S, H, hd, scale)
    e_tc, e_mma = rel_fro(outs[0].double(), ref), rel_fro(outs[1].double(), ref)
    print(f"spatial attention variant {variant} F={n_frames} S={S} H={H} peaked={peaked}: tcgen05 rel {e_tc:.3e}, mma.sync rel {e_mma:.3e}, "
          f"max abs diff between kernels {float((outs[0].float() - outs[1].float()).abs().max()):.3e}")
    assert torch.isfinite(outs[0].float()).all()
    assert e_tc < 6e-3          # bf16 P and bf16 output rounding
    assert e_tc < 1.5 * e_mma + 1e-3


@pytest.mark.parametrize("fp16", [False, True])
@pytest.mark.parametrize("n_frames,H,hd,peaked", [(3, 8, 32, False), (40, 8, 32, True), (2, 2, 32, True), (3, 8, 64, True)])
def test_spatial_attention_tcgen05_head_dim_32_and_fp16(lib, n_frames, H, hd, peaked, fp16):
    """round 2: (a) head_dim 32 (the in-tree 35M config) on the persistent tcgen05 kernel - 64-wide head-PAIR boxes,
    the head's own two 32-byte K slices for Q K^T, its own 32 accumulator columns in the epilogue: a wrong slice / column
    mapping mixes the two heads of a pair and cannot pass; 40 x 8 items > 148 CTAs exercises the ring; (b) IEEE fp16
    data (kernel | 0x100) through the same kernels: P and the output are rounded to 11 bits instead of 8."""
    L, _l = lib
    S, d = 256, H * hd
    dt = torch.float16 if fp16 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(300 + n_frames + H + hd)
    qkv = torch.randn(n_frames * S, 3 * d, device="cuda", generator=g)
    if peaked:
        qkv[:, :2 * d] *= 3.0
    qkv = qkv.to(dt).contiguous()
    scale = 1.0 / math.sqrt(hd)
    flag = 0x100 if fp16 else 0
    outs = []
    for kernel in (0, 1):                                   # 0: tcgen05, 1: mma.sync kernel it replaces
        out = torch.full((n_frames * S, d), float("nan"), device="cuda", dtype=dt)
        _l.check(L.gn_spatial_attention(P(qkv), P(out), n_frames, S, H, hd, scale, kernel | flag, None))
        torch.cuda.synchronize()
        outs.append(out)
    ref = _spatial_ref(qkv, n_frames, S, H, hd, scale)
    e_tc, e_mma = rel_fro(outs[0].double(), ref), rel_fro(outs[1].double(), ref)
    print(f"spatial attention hd={hd} fp16={fp16} F={n_frames} H={H} peaked={peaked}: tcgen05 rel {e_tc:.3e}, "
          f"mma.sync rel {e_mma:.3e}")
    assert torch.isfinite(outs[0].float()).all()
    assert e_tc < (8e-4 if fp16 else 6e-3)
    assert e_tc < 1.5 * e_mma + (2e-4 if fp16 else 1e-3)


@pytest.mark.parametrize("epi,N,K", [(0, 1536, 512), (1, 2048, 512), (2, 512, 2048), (2, 512, 512)])
def test_linear_fp16_operands(lib, epi, N, K):
    """the tcgen05 linear kernel with IEEE fp16 operands / outputs (in_bf16 = out_bf16 = 2) against float64."""
    L, _l = lib
    M = 1000                                                 # ragged last tile
    g = torch.Generator(device="cuda").manual_seed(7 + epi + N)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g) if epi == 2 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi == 2 else torch.float16)
    out2 = torch.empty(M, N, device="cuda", dtype=torch.float16) if epi == 2 else None
    _l.check(L.gn_linear_forward(P(a), P(w), P(b), P(r), P(out), P(out2), M, N, K, epi, 2, 0 if epi == 2 else 2, 0, None))
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + b.double()
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = ref + r.double()
    err = rel_fro(out.double(), ref)
    print(f"fp16 linear epi={epi} N={N} K={K}: rel {err:.3e}")
    assert err < (1e-5 if epi == 2 else 4e-4)                # fp32 residual output vs fp16-rounded output
    if out2 is not None:
        assert rel_fro(out2.double(), ref) < 4e-4
