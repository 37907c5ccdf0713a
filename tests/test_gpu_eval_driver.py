"""evaluate.py driver on the GPU: sharded teacher-forced CE/acc equals the CPU oracle; with >= 2 GPUs the same
result comes out of a 2-rank NCCL job (one all-reduce of 4 doubles)."""
import importlib
import os
import socket
import subprocess
import sys

import pytest
import torch

from helpers import O, ROOT, build_b200_model, golden_cfg, golden_sd, load_golden

pytestmark = pytest.mark.gpu


def _oracle_metrics(sd, cfg, clips, steps, noise_seed):
    B = clips.shape[0]
    # same per-clip noise derivation as evaluate.b200_backend: generator seeded with (noise_seed + global clip index)
    noise = torch.stack([torch.rand(cfg.T - 1, steps - 1, cfg.S, generator=torch.Generator().manual_seed(noise_seed + i))
                         for i in range(B)], dim=2)                                  # [T-1, K-1, B, S]
    return O.teacher_forced_metrics(sd, cfg, clips.reshape(B, -1), steps, noise)


def test_single_gpu_driver_matches_oracle():
    ev = importlib.import_module("1xgpt_b200.evaluate")
    z = load_golden("tiny_preln")
    kw, sd = golden_cfg(z), golden_sd(z)
    cfg = O.OracleConfig(**kw)
    clips = O.synthetic_clips(cfg, 5, seed=77).reshape(5, -1)
    loss, acc, _ = _oracle_metrics(sd, cfg, clips, 2, 1234)
    for kv in (False, True):
        m = build_b200_model(kw, sd, precision="fp32", kv_cache=kv)
        res = ev.evaluate_clips(clips, ev.b200_backend(m, maskgit_steps=2, noise_seed=1234), batch_size=2,
                                acc_device=m.device, rank=0, world=1)
        assert res["tokens"] == 5 * (cfg.T - 1) * cfg.S
        assert abs(res["loss"] - loss) < 1e-4
        assert abs(res["acc"] - acc) < 1e-9


WORKER = r'''
import importlib, json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
from helpers import O, build_b200_model, golden_cfg, golden_sd, load_golden
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ev = importlib.import_module("1xgpt_b200.evaluate")
z = load_golden("tiny_preln"); kw, sd = golden_cfg(z), golden_sd(z); cfg = O.OracleConfig(**kw)
clips = O.synthetic_clips(cfg, 5, seed=77).reshape(5, -1)
pkg = importlib.import_module("1xgpt_b200")
m = pkg.STMaskGIT(pkg.GenieConfig(**kw), precision="fp32", kv_cache=True); m.load_state_dict(sd); m = m.to(f"cuda:{rank}")
res = ev.evaluate_clips(clips, ev.b200_backend(m, maskgit_steps=2, noise_seed=1234), batch_size=2, acc_device=m.device)
print("RESULT", json.dumps(res), flush=True)
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_driver_matches_oracle(tmp_path):
    z = load_golden("tiny_preln")
    kw, sd = golden_cfg(z), golden_sd(z)
    cfg = O.OracleConfig(**kw)
    clips = O.synthetic_clips(cfg, 5, seed=77).reshape(5, -1)
    loss, acc, _ = _oracle_metrics(sd, cfg, clips, 2, 1234)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, REPO_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    results = [json.loads(l.split("RESULT ", 1)[1]) for l in out.stdout.splitlines() if "RESULT " in l]
    assert len(results) == 2
    for r in results:
        assert r["world"] == 2 and r["tokens"] == 5 * (cfg.T - 1) * cfg.S
        assert abs(r["loss"] - loss) < 1e-4 and abs(r["acc"] - acc) < 1e-9
    assert sorted(r["local_clips"] for r in results) == [2, 3]
