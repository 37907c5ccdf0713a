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
