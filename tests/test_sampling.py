"""temperature > 0 branch of maskgit_generate (st_mask_git.py:182-187): what the reference samples from, pinned by
histograms the unmodified reference produced (tests/golden/tiny_sampling.npz, make_golden_sampling.py), and the
oracle's inverse-CDF sampler that the CUDA kernel mirrors."""
import numpy as np
import torch

from helpers import O, golden_cfg, load_golden


def _pooled_chi2(counts, probs, n):
    """Pearson chi-square of one histogram against n*probs; bins with expectation < 5 are pooled into one."""
    exp = probs * n
    small = exp < 5
    c = np.concatenate([counts[~small], [counts[small].sum()]]) if small.any() else counts
    e = np.concatenate([exp[~small], [exp[small].sum()]]) if small.any() else exp
    keep = e > 0
    c, e = c[keep], e[keep]
    return float(((c - e) ** 2 / e).sum()), len(e) - 1


def _zscore(counts_PNV, probs_PNV, n):
    chi2, dof = 0.0, 0
    for p in range(counts_PNV.shape[0]):
        for i in range(counts_PNV.shape[1]):
            c, d = _pooled_chi2(counts_PNV[p, i].astype(np.float64), probs_PNV[p, i], n)
            chi2 += c
            dof += d
    return (chi2 - dof) / np.sqrt(2.0 * dof), dof


def _golden_probs(z):
    l0 = torch.from_numpy(z["logits0"])                       # [B, V, NV, H, W]
    B, V, NV = l0.shape[:3]
    probs = torch.softmax(l0.double(), dim=1).reshape(B, V, NV, -1)
    return probs.permute(0, 3, 2, 1).reshape(-1, NV, V).numpy()   # [B*S, NV, V]


def test_reference_histograms_follow_softmax_at_every_temperature():
    z = load_golden("tiny_sampling")
    probs = _golden_probs(z)
    n = int(z["n_draws"])
    assert z["counts"].sum(axis=-1).min() == n and z["counts"].sum(axis=-1).max() == n
    for ti, temp in enumerate(z["temps"]):
        zs, dof = _zscore(z["counts"][ti], probs, n)
        print(f"reference, temperature {temp}: chi2 z-score {zs:+.2f} over {dof} dof")
        assert abs(zs) < 5.0
    # a sampler that actually applied the temperature (softmax(logits / T)) is rejected by the same statistic
    l0 = torch.from_numpy(z["logits0"]).double()
    B, V, NV = l0.shape[:3]
    hot = torch.softmax(l0 / 3.0, dim=1).reshape(B, V, NV, -1).permute(0, 3, 2, 1).reshape(-1, NV, V).numpy()
    zs, _ = _zscore(z["counts"][2], hot, n)
    assert zs > 50.0


def test_oracle_inverse_cdf_matches_reference_distribution():
    z = load_golden("tiny_sampling")
    probs = _golden_probs(z)                                  # [P, NV, V]
    P, NV, V = probs.shape
    n = int(z["n_draws"])
    g = torch.Generator().manual_seed(5)
    counts = np.zeros((P, NV, V), dtype=np.int64)
    pt = torch.from_numpy(probs).permute(2, 0, 1).reshape(1, V, P, NV)      # categorical_icdf wants [B, V, ...]
    for _ in range(n):
        u = torch.rand(1, P, NV, generator=g)
        s = O.categorical_icdf(pt, u).reshape(P, NV).numpy()
        for i in range(NV):
            counts[np.arange(P), i, s[:, i]] += 1
    zs, dof = _zscore(counts, probs, n)
    print(f"oracle inverse CDF: chi2 z-score {zs:+.2f} over {dof} dof")
    assert abs(zs) < 5.0
    # two-sample check against the reference's own draws (temperature 1.0): same pooled statistic on the difference
    ref = z["counts"][1].astype(np.float64)
    diff = ((counts - ref) ** 2 / np.maximum(counts + ref, 1.0))[(counts + ref) >= 10].sum()
    k = int(((counts + ref) >= 10).sum())
    assert abs(diff - k) / np.sqrt(2.0 * k) < 5.0


def test_inverse_cdf_edges_and_determinism():
    p = torch.tensor([0.0, 0.25, 0.0, 0.5, 0.25]).reshape(1, 5, 1)
    f = lambda u: int(O.categorical_icdf(p, torch.tensor([[u]])))
    assert f(0.0) == 1                     # first index with mass
    assert f(0.2499) == 1 and f(0.25) == 3 and f(0.7499) == 3 and f(0.75) == 4
    assert f(0.999999) == 4
    # unnormalised input: Categorical divides by the sum, the inverse CDF scales the target instead
    assert int(O.categorical_icdf(p * 7.0, torch.tensor([[0.5]]))) == 3
    d = torch.distributions.Categorical(probs=torch.tensor([0.1, 0.2, 0.7]) / 3.0)
    assert torch.allclose(d.probs, torch.tensor([0.1, 0.2, 0.7]))       # the temperature cancels in the reference


def test_oracle_maskgit_generate_with_temperature_is_reproducible():
    z = load_golden("tiny_preln")
    kw = golden_cfg(z)
    cfg = O.OracleConfig(**kw)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    prompt = torch.from_numpy(z["prompt"])
    B = prompt.shape[0]
    noise = torch.from_numpy(z["noise"])
    u = torch.rand(3, B, cfg.S, cfg.num_factored_vocabs, generator=torch.Generator().manual_seed(9))
    a, l0 = O.maskgit_generate(sd, cfg, prompt.clone(), 2, 3, temperature=1.0, noise=noise, uniform=u)
    b, _ = O.maskgit_generate(sd, cfg, prompt.clone(), 2, 3, temperature=0.3, noise=noise, uniform=u)
    assert torch.equal(a, b)                                           # temperature value is irrelevant
    g, _ = O.maskgit_generate(sd, cfg, prompt.clone(), 2, 3, temperature=0.0, noise=noise)
    assert torch.equal(g, torch.from_numpy(z["samples"]))              # greedy branch untouched
    assert not torch.equal(a, g)
    assert torch.equal(l0, torch.from_numpy(z["logits0"]))
