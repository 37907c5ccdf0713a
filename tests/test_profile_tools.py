"""The scripts that turn ncu output into the committed profiles/ summaries, run on the committed raw files
(no GPU): bench.py's `roofline.traffic` and the launch-share table must be reproducible from them."""
import gzip
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_insitu_traffic_json_reproducible(tmp_path):
    out = tmp_path / "t.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_insitu_traffic.py"),
                    os.path.join(PROF, "r01_gemm_insitu_final_ncu_full_raw.csv"), str(out), "test"], check=True,
                   capture_output=True)
    new = json.load(open(out))
    ref = json.load(open(os.path.join(PROF, "r01_gemm_insitu_final_traffic.json")))
    assert len(new["launches"]) == len(ref["launches"]) == 12
    assert abs(new["avg_dram_bytes_per_gemm_launch"] - ref["avg_dram_bytes_per_gemm_launch"]) < 1.0
    assert all("gemm_tcgen05_kernel" in l["kernel"] for l in new["launches"])


def test_launch_share_table_reproducible(tmp_path):
    csv_path = tmp_path / "launches.csv"
    with gzip.open(os.path.join(PROF, "r01_launches_final.csv.gz"), "rt") as f:
        csv_path.write_text(f.read())
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), str(csv_path), "t"],
                       check=True, capture_output=True, text=True)
    committed = open(os.path.join(PROF, "r01_launch_shares_final.md")).read()
    rows = [ln for ln in r.stdout.splitlines() if ln.startswith("| `")]
    assert len(rows) >= 10
    for ln in rows[:8]:                      # the kernels that carry the step
        assert ln in committed, ln
    assert "tcgen05 GEMM variants together 66.1%" in r.stdout


def test_round2_ncu_evidence_reproducible(tmp_path):
    """round 2: the per-kernel table (HBM GB/s and tensor-pipe activity of EVERY kernel family, profiles/r02_ncu_kernels)
    and bench.py's roofline.traffic source are regenerated from the committed raw ncu pages."""
    raws = sorted(os.path.join(PROF, "r02_ncu", f) for f in os.listdir(os.path.join(PROF, "r02_ncu")) if f.endswith("_raw.csv.gz"))
    assert len(raws) >= 8
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), str(tmp_path / "k")] + raws, check=True,
                   capture_output=True)
    new = json.load(open(tmp_path / "k.json"))
    ref = json.load(open(os.path.join(PROF, "r02_ncu_kernels.json")))
    assert [k["kernel"] for k in new["kernels"]] == [k["kernel"] for k in ref["kernels"]]
    names = " ".join(k["kernel"] for k in new["kernels"])
    for fam in ("gemm_tcgen05_kernel", "spatial_attn_tc_persistent_kernel", "temporal_attn_v2_kernel", "prep_kernel",
                "embed_kernel", "readout_sample_kernel", "remask_kernel", "ce_kernel", "count_equal_kernel",
                "gn_partial_kernel", "gn_apply_swish_kernel", "stem_conv_kernel", "vq_head_kernel"):
        assert fam in names, fam
    for k in new["kernels"]:
        assert k["avg_us"] > 0 and k["hbm_gbs"] >= 0
    with gzip.open(os.path.join(PROF, "r02_ncu", "r02_gemm_insitu_raw.csv.gz"), "rt") as f:
        (tmp_path / "g.csv").write_text(f.read())
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_insitu_traffic.py"), str(tmp_path / "g.csv"),
                    str(tmp_path / "t.json"), "test"], check=True, capture_output=True)
    t_new, t_ref = json.load(open(tmp_path / "t.json")), json.load(open(os.path.join(PROF, "r02_gemm_insitu_traffic.json")))
    assert abs(t_new["avg_dram_bytes_per_gemm_launch"] - t_ref["avg_dram_bytes_per_gemm_launch"]) < 1.0


def test_round2b_magvit_launch_lists_reproducible(tmp_path):
    """second round-2 session: the three MAGVIT2 launch-share tables (32 images per pass; + CTA-pair conv tiles and the
    mma.sync output conv; + the fused GroupNorm kernel) are regenerated from the committed raw ncu launch lists."""
    for raw, md in (("r02b_vq_launches_per32", "r02b_launch_shares_magvit_per32"),
                    ("r02b_vq_launches_pair_mma", "r02b_launch_shares_magvit_pair_mma"),
                    ("r02b_vq_launches_gn_fused", "r02b_launch_shares_magvit_gn_fused")):
        csv_path = tmp_path / (raw + ".csv")
        with gzip.open(os.path.join(PROF, "r02b_ncu", raw + ".csv.gz"), "rt") as f:
            csv_path.write_text(f.read())
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), str(csv_path), "t"],
                           check=True, capture_output=True, text=True)
        committed = open(os.path.join(PROF, md + ".md")).read()
        rows = [ln for ln in r.stdout.splitlines() if ln.startswith("| `")]
        assert len(rows) >= 12
        for ln in rows:
            assert ln in committed, (md, ln)


def test_round2b_ncu_pages_reproducible(tmp_path):
    """--set full pages of the two MAGVIT2 kernels that replaced round-2 ones (8-row stem conv, mma.sync output conv)."""
    raw = os.path.join(PROF, "r02b_ncu", "r02b_vq_new_raw.csv.gz")
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), str(tmp_path / "k"), raw], check=True,
                   capture_output=True)
    new = json.load(open(tmp_path / "k.json"))
    ref = json.load(open(os.path.join(PROF, "r02b_ncu_kernels.json")))
    assert [k["kernel"] for k in new["kernels"]] == [k["kernel"] for k in ref["kernels"]]
    names = " ".join(k["kernel"] for k in new["kernels"])
    assert "stem_conv_kernel<4, 8>" in names and "out_conv_mma_kernel" in names
    for k in new["kernels"]:
        assert k["avg_us"] > 0 and k["hbm_gbs"] > 0
