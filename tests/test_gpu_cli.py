"""End-to-end CLI tests: `cli_generate.main` / `cli_evaluate.main` (twins of genie/generate.py:62-116 and
genie/evaluate.py:146-191) run on the GPU against
  * a dataset directory in the reference's on-disk format (video.bin uint32 [N,s,s], segment_ids.bin, metadata.json)
    that this file synthesises, and
  * tests/golden/ckpt_tiny: a checkpoint directory written by the REFERENCE's own save_pretrained
    (tests/golden/make_golden_r2.py),
and are compared with what the reference's generate / evaluate loops produced from the same two directories
(tests/golden/ckpt_tiny_expected.npz)."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT, load_golden

CKPT = os.path.join(ROOT, "tests", "golden", "ckpt_tiny")


def synthetic_dataset(side, n_frames=700, boundary=300, seed=5, vocab=262144):
    """same stream as tests/golden/make_golden_r2.py:synthetic_dataset"""
    g = np.random.default_rng(seed)
    video = g.integers(0, vocab, size=(n_frames, side, side), dtype=np.uint32)
    seg = np.zeros(n_frames, dtype=np.int32)
    seg[boundary:] = 1
    return video, seg


def write_synthetic_dataset(path):
    os.makedirs(path, exist_ok=True)
    video, seg = synthetic_dataset(4)
    video.tofile(os.path.join(path, "video.bin"))
    seg.tofile(os.path.join(path, "segment_ids.bin"))
    with open(os.path.join(path, "metadata.json"), "w") as f:
        json.dump({"num_images": int(video.shape[0]), "s": 4, "vocab_size": 262144, "hz": 2, "token_dtype": "uint32"}, f)
    return str(path)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_cli_generate_main_matches_reference(tmp_path, precision):
    cli = importlib.import_module("1xgpt_b200.cli_generate")
    z = load_golden("ckpt_tiny_expected")
    data = write_synthetic_dataset(tmp_path / "val")
    out = tmp_path / "gen"
    cli.main(["--checkpoint_dir", CKPT, "--val_data_dir", data, "--output_dir", str(out), "--maskgit_steps", "1",
              "--precision", precision])
    meta = json.load(open(out / "metadata.json"))
    assert (meta["num_images"], meta["h"], meta["w"], meta["t"]) == (24, 4, 4, 16)     # generate.py:104-113
    assert meta["token_dtype"] == "uint32" and meta["maskgit_steps"] == 1
    frames = np.fromfile(out / "video.bin", dtype=np.uint32).reshape(24, 4, 4)
    ex, gen = z["example"][0], z["generated"][0]
    assert np.array_equal(frames[:8], ex[:8])                  # prompt frames
    assert np.array_equal(frames[16:], ex[8:])                 # ground truth
    agree = float((frames[8:16] == gen[8:]).mean())
    print(f"cli_generate {precision}: generated-token agreement with the reference {agree:.4f}")
    if precision == "fp32":
        assert np.array_equal(frames[8:16], gen[8:])           # bit-exact ids in the exact mode
    else:
        assert agree >= 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_cli_evaluate_main_matches_reference(tmp_path, capsys, precision):
    cli = importlib.import_module("1xgpt_b200.cli_evaluate")
    z = load_golden("ckpt_tiny_expected")
    data = write_synthetic_dataset(tmp_path / "val")
    cli.main(["--checkpoint_dir", CKPT, "--val_data_dir", data, "--maskgit_steps", "1", "--max_examples", "3",
              "--batch_size", "2", "--precision", precision])
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert int(res["tokens"]) == 3 * 15 * 16
    tol = 1e-4 if precision == "fp32" else 2e-3
    assert abs(float(res["loss"]) - float(z["eval_loss"])) < tol * float(z["eval_loss"])
    if precision == "fp32":
        assert abs(float(res["acc"]) - float(z["eval_acc"])) < 1e-9


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cli_evaluate_two_ranks_nccl(tmp_path):
    """the same evaluation sharded over 2 ranks (one NCCL all-reduce) prints the same loss"""
    z = load_golden("ckpt_tiny_expected")
    data = write_synthetic_dataset(tmp_path / "val")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29577", "-m", "1xgpt_b200.cli_evaluate", "--checkpoint_dir", CKPT,
           "--val_data_dir", data, "--maskgit_steps", "1", "--max_examples", "3", "--batch_size", "2", "--precision",
           "fp32"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert int(res["tokens"]) == 3 * 15 * 16 and int(res["world"]) == 2
    assert abs(float(res["loss"]) - float(z["eval_loss"])) < 1e-4 * float(z["eval_loss"])
