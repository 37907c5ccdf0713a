"""world_size-2 gloo test of the batch-sharded evaluation driver (1xgpt_b200/evaluate.py): sharding + the one
all-reduce(sum) of the 4-scalar accumulator.  The per-batch backend here is the CPU ORACLE (checker), so the
test also shows that the sharded result equals the unsharded oracle result."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, golden_cfg, golden_sd, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_backend(sd, cfg, steps):
    def fn(batch, first_index):
        B = batch.shape[0]
        noise = torch.stack([torch.stack([O.tie_free_noise(steps, 1, cfg.S, seed=1000 + first_index + i)[:, 0]
                                          for i in range(B)], dim=1) for _ in range(cfg.T - 1)])  # [T-1,K-1,B,S]
        samples, fl = O.predict_zframe_logits(sd, cfg, batch.reshape(B, -1), steps, 0.0, noise)
        V, NV = cfg.factored_vocab_size, cfg.num_factored_vocabs
        gt = batch.reshape(B, cfg.T, cfg.hw, cfg.hw)[:, 1:]
        labels = O.factorize_labels(gt, NV, V)
        ce = torch.nn.functional.cross_entropy(fl.double(), labels, reduction="none").sum(dim=1)
        ok = (fl.argmax(dim=1) == labels).all(dim=1)
        return torch.tensor([float(ce.sum()), float(ce.numel()), float(ok.sum()), float((gt == samples).sum())],
                            dtype=torch.float64)
    return fn


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    ev = importlib.import_module("1xgpt_b200.evaluate")
    z = load_golden("tiny_preln")
    kw, sd = golden_cfg(z), golden_sd(z)
    cfg = O.OracleConfig(**kw)
    clips = O.synthetic_clips(cfg, 5, seed=77).reshape(5, -1)          # 5 clips over 2 ranks: 3 + 2
    out = ev.evaluate_clips(clips, _oracle_backend(sd, cfg, 2), batch_size=2, acc_device="cpu")
    q.put((rank, out))
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    ev = importlib.import_module("1xgpt_b200.evaluate")
    for n in (0, 1, 5, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [ev.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_eval_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0]["local_clips"] == 3 and res[1]["local_clips"] == 2
    for k in ("loss", "acc", "argmax_acc", "tokens"):
        assert res[0][k] == res[1][k]                     # every rank holds the reduced result
    # single-process reference over all 5 clips
    ev = importlib.import_module("1xgpt_b200.evaluate")
    z = load_golden("tiny_preln")
    kw, sd = golden_cfg(z), golden_sd(z)
    cfg = O.OracleConfig(**kw)
    clips = O.synthetic_clips(cfg, 5, seed=77).reshape(5, -1)
    one = ev.evaluate_clips(clips, _oracle_backend(sd, cfg, 2), batch_size=5, acc_device="cpu", rank=0, world=1)
    assert res[0]["tokens"] == one["tokens"] == 5 * (cfg.T - 1) * cfg.S
    assert abs(res[0]["loss"] - one["loss"]) < 1e-9
    assert res[0]["acc"] == one["acc"]
