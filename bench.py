#!/usr/bin/env python3
"""bench.py — generated frames/s of the GENIE_138M MaskGIT-2 generate path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode cached|dense] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): GENIE_138M shape (L32 d512 h8, pre-LN, 2x512 factored vocab), random-init
weights, synthetic clips; one STEP = one `generate` pass over a batch of 64 clips per GPU: 8 prompt frames ->
8 generated frames per clip, MaskGIT 2 steps, temperature 0  =>  512 generated frames per step per GPU.
Clips shard over ranks with no data-path collective (weak scaling); the timed region is bracketed by a
barrier + cuda synchronize on both sides, timed with CUDA events, max over ranks.

JSON line keys beyond the base contract: roofline (tcgen05 linear kernel, timed live with CUDA events around
every launch in the timed region), cpu_baseline (the oracle port on the host cores, bounded sample), e2e (same
metric through gn_generate_host with pinned HOST buffers: H2D of tokens+noise and D2H of tokens inside the
timed region), clocks (nvidia-smi sampled during the timed region), gpu_launches.

`--impl reference` times the reference's CPU implementation of the path (the oracle port: the reference is a
Python package and cannot travel to the GPU box; oracle/genie_oracle.py is its bit-exact restatement, pinned by
tests/golden) with all host threads on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL_KW = dict(num_layers=32, num_heads=8, d_model=512, T=16, S=256, image_vocab_size=262144, use_mup=False,
                num_factored_vocabs=2, qkv_bias=False, proj_bias=True, qk_norm=False, mlp_ratio=4.0, mlp_bias=True)
T_PROMPT, MASKGIT_STEPS = 8, 2
METRIC = "generated frames/sec (GENIE_138M, 16x256-token clips, MaskGIT-2)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synth_state_dict(seed=0):
    """N(0, 0.02) weights with the reference's state_dict keys (st_mask_git.py:281-296 style init)."""
    from oracle import genie_oracle as O
    cfg = O.OracleConfig(**MODEL_KW)
    return cfg, O.init_state_dict(cfg, seed=seed, bias_std=0.02)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = max(smax, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pick_cpu_threads():
    """torch's CPU kernels stop scaling (and thrash) far below the core count of a 2-socket GPU host, so "all the
    host threads it can use" is found by a 2-second calibration: one 4-layer forward of the same shape at
    8/16/32/64/all threads, keep the fastest."""
    from oracle import genie_oracle as O
    ncpu = os.cpu_count() or 1
    kw = dict(MODEL_KW, num_layers=4)
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=1)
    ids = O.synthetic_clips(cfg, 1, seed=2)
    best, best_t = 1, float("inf")
    for n in sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu} | {min(ncpu, 8)}):
        torch.set_num_threads(n)
        with torch.no_grad():
            O.compute_logits(sd, cfg, ids)  # warm
            t0 = time.perf_counter()
            O.compute_logits(sd, cfg, ids)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    return best


def cpu_generate_sample(cfg, sd, batch, reps, threads):
    """Oracle port: maskgit_generate of ONE frame (out_t = 8, K = 2) for `batch` clips = 2 full-window forwards
    per clip, exactly the reference's per-frame work (st_mask_git.py:163,169).  frames/s = batch / seconds."""
    from oracle import genie_oracle as O
    torch.set_num_threads(threads)
    ids = O.synthetic_clips(cfg, batch, seed=4321)
    ids[:, T_PROMPT:] = cfg.mask_token_id
    noise = O.tie_free_noise(MASKGIT_STEPS, batch, cfg.S, seed=5)
    times = []
    for r in range(reps + 1):
        p = ids.clone()
        t0 = time.perf_counter()
        with torch.no_grad():
            O.maskgit_generate(sd, cfg, p, T_PROMPT, MASKGIT_STEPS, 0.0, noise=noise)
        dt = time.perf_counter() - t0
        if r > 0 or reps == 0:
            times.append(dt)
    return batch / statistics.median(times), times


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = pick_cpu_threads()
    cfg, sd = synth_state_dict()
    batch = 2
    times = []
    torch.set_num_threads(threads)
    from oracle import genie_oracle as O
    ids = O.synthetic_clips(cfg, batch, seed=4321)
    ids[:, T_PROMPT:] = cfg.mask_token_id
    noise = O.tie_free_noise(MASKGIT_STEPS, batch, cfg.S, seed=5)
    for i in range(args.warmup + args.steps):
        p = ids.clone()
        t0 = time.perf_counter()
        with torch.no_grad():
            O.maskgit_generate(sd, cfg, p, T_PROMPT, MASKGIT_STEPS, 0.0, noise=noise)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = batch * len(times) / total
    sample = (f"oracle port (bit-exact restatement of the reference modules), fp32, {threads} threads (fastest "
              f"of 8/16/32/64/{os.cpu_count()} in a short calibration): "
              f"maskgit_generate of 1 frame (out_t=8, K=2 => 2 full 16x256 forwards) for {batch} clips per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "GENIE_138M generate.py MaskGIT 2-step temperature=0, batch 64 clips per GPU "
                               "(BASELINE.json configs[1])",
                   "sample": "bounded CPU sample of that workload: 1 generated frame for 2 clips per step",
                   "clips_per_step": batch, "frames_per_step": batch},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="cached", choices=["cached", "dense"],
                    help="cached: temporal K/V cache + causal frame trimming (identical tokens); dense: recompute the "
                         "full 16-frame window every MaskGIT step like the reference")
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--chunk-tokens", type=int, default=32768)
    ap.add_argument("--fold-ln", action="store_true", help="LayerNorm folded into the QKV/fc1 GEMM epilogues")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--lanes", type=int, default=0,
                    help="concurrent streams the clips of a MaskGIT step are dealt to (0 = library default, 1 = single)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary (other mode) measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # a non-default stream, so the library can capture / replay its per-chunk CUDA graphs
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    pkg = importlib.import_module("1xgpt_b200")
    lib = pkg._lib.load()

    cfg, sd = synth_state_dict()
    B, T, S = args.batch, cfg.T, cfg.S
    n_new = T - T_PROMPT

    def make_model(mode):
        m = pkg.STMaskGIT(pkg.GenieConfig(**MODEL_KW), precision="bf16", kv_cache=(mode == "cached"),
                          chunk_tokens=args.chunk_tokens, fold_ln=args.fold_ln, cuda_graphs=not args.no_graphs, lanes=args.lanes)
        m.load_state_dict(sd)
        return m.to(dev)

    model = make_model(args.mode)
    h = model._handle()
    g = torch.Generator().manual_seed(1234 + rank)
    clips = torch.randint(0, cfg.image_vocab_size, (B, T, S), generator=g, dtype=torch.int32)
    noise = torch.stack([torch.stack([torch.stack([torch.randperm(S, generator=g).float() / S for _ in range(B)])
                                      for _ in range(MASKGIT_STEPS - 1)]) for _ in range(n_new)])  # [8, K-1, B, S]
    clips_pin, noise_pin = clips.pin_memory(), noise.contiguous().pin_memory()
    clips_dev, noise_dev = clips.to(dev), noise.to(dev).contiguous()
    work = torch.empty_like(clips_dev)
    stream = torch.cuda.current_stream(dev)
    sptr = C.c_void_p(stream.cuda_stream)

    def step_device(mh):
        work.copy_(clips_dev)
        pkg._lib.check(lib.gn_generate(mh.ptr, C.c_void_p(work.data_ptr()), B, T_PROMPT, MASKGIT_STEPS, 0.0, 0,
                                       C.c_void_p(noise_dev.data_ptr()), None, None, sptr))

    host_buf = torch.empty_like(clips_pin).pin_memory()

    def step_host(mh):
        host_buf.copy_(clips_pin)
        pkg._lib.check(lib.gn_generate_host(mh.ptr, C.c_void_p(host_buf.data_ptr()), B, T_PROMPT, MASKGIT_STEPS, 0.0,
                                            0, C.c_void_p(noise_pin.data_ptr()), None, sptr))
        return int(host_buf[0, T - 1, 0])  # touch the result (already synchronised by the call)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, mh, steps, warmup, profile=False):
        for _ in range(warmup):
            fn(mh)
        barrier()
        launches0 = lib.gn_kernel_launches()
        model.reset_counters() if mh is h else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            lib.gn_profile_begin()
        e0.record(stream)
        t_host = time.perf_counter()
        for _ in range(steps):
            fn(mh)
        timed.host_enqueue_ms = 1e3 * (time.perf_counter() - t_host) / steps   # CPU time to enqueue one step
        e1.record(stream)
        barrier()
        prof = None
        if profile:
            out = (C.c_double * 15)()
            pkg._lib.check(lib.gn_profile_end(out))
            prof = list(out)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, lib.gn_kernel_launches() - launches0, prof

    frames_per_step = B * n_new * world
    # ---- device-resident throughput (headline `value`)
    with ClockSampler(local) as cs:
        ms, launches, _ = timed(step_device, h, args.steps, args.warmup)
    clocks = cs.summary()
    host_enqueue_ms = timed.host_enqueue_ms
    flops_exec = model.flops_executed()
    value = frames_per_step * args.steps / (ms / 1e3)
    # ---- the same K steps again with every tcgen05 GEMM launch bracketed by CUDA events on its stream (the event
    #      records cost a few percent of the step, so they are kept out of the headline region)
    ms_prof, _, prof = timed(step_device, h, args.steps, 1, profile=True)
    # ---- end-to-end through the host-buffer C-ABI call
    ms_e2e, _, _ = timed(step_host, h, args.steps, max(1, args.warmup // 2))
    e2e = frames_per_step * args.steps / (ms_e2e / 1e3)

    peaks, peak_src = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    cats = ["gemm_store", "gemm_gelu", "gemm_resid", "prep_ln", "spatial_attn", "temporal_attn", "other"]
    by_cat = {c: {"ms_per_step": prof[2 * i] / args.steps, "launches_per_step": prof[2 * i + 1] / args.steps}
              for i, c in enumerate(cats)}
    gemm_ms = prof[0] + prof[2] + prof[4]
    gemm_launches = prof[1] + prof[3] + prof[5]
    gemm_flops = prof[14]
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    dense_flops_step = model.flops_per_clip_forward() * B * n_new * MASKGIT_STEPS  # reference-equivalent FLOPs / step / GPU

    # DRAM bytes per GEMM launch from the committed ncu --set full capture of the same kernel in situ (profiles/)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_insitu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("avg_dram_bytes_per_gemm_launch")

    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        other = "dense" if args.mode == "cached" else "cached"
        del model
        torch.cuda.empty_cache()
        m2 = make_model(other)
        h2 = m2._handle()
        n2 = max(1, min(args.steps, 2))
        ms2, _, _ = timed(step_device, h2, n2, 1)
        secondary = {"mode": other, "value": B * n_new * n2 / (ms2 / 1e3), "unit": "frames/s", "steps": n2,
                     "ms_per_step": ms2 / n2}
        model = m2

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = pick_cpu_threads()
        v, times = cpu_generate_sample(cfg, sd, batch=2, reps=2, threads=threads)
        cpu = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"oracle port fp32, {threads} threads (fastest of 8/16/32/64/{os.cpu_count()} in a short "
                         f"calibration): maskgit_generate of 1 frame (K=2, full 16x256 window) for 2 clips, "
                         f"1 warm-up + 2 reps, median {statistics.median(times):.2f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "GENIE_138M generate.py MaskGIT 2-step temperature=0, batch 64 clips per GPU "
                                   "(BASELINE.json configs[1])",
                       "clips_per_gpu": B, "prompt_frames": T_PROMPT, "new_frames": n_new,
                       "maskgit_steps": MASKGIT_STEPS, "frames_per_step": frames_per_step, "mode": args.mode,
                       "kv_cache": args.mode == "cached", "lanes": args.lanes, "precision": "bf16 operands, fp32 accumulate/residual",
                       "cache_hygiene": "per-step working set (275 MB weights + >10 GB activations/KV) exceeds the "
                                        "126 MB L2; no L2 flush needed",
                       "parallelism": f"dp{world} (clips sharded, no collective)"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic,
                         "traffic_note": "avg dram__bytes_read+write per tcgen05 GEMM launch, ncu --set full in situ at "
                                         "M=32256 (profiles/r01_gemm_insitu_traffic.json); algorithmic bytes per "
                                         "launch (A + residual + outputs) avg 171 MB at that M",
                         "kernel": "gemm_tcgen05_kernel (all linear layers)", "launches_timed": int(gemm_launches),
                         "kernel_ms_per_step": gemm_ms / args.steps, "kernel_share_of_step": gemm_ms / ms_prof,
                         "profiled_ms_per_step": ms_prof / args.steps, "kernel_ms_by_category": by_cat,
                         "peak_source": f"{peak_src} bf16_tflops_sustained"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(clips_pin.numel() * 4 + noise_pin.numel() * 4),
                    "d2h_bytes_per_step": int(clips_pin.numel() * 4)},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
            "flops": {"executed_per_step_per_gpu": flops_exec / args.steps,
                      "dense_reference_equivalent_per_step_per_gpu": dense_flops_step,
                      "model_tflops_executed": flops_exec / (ms / 1e3) / 1e12,
                      "dense_equivalent_tflops": dense_flops_step * args.steps / (ms / 1e3) / 1e12},
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
