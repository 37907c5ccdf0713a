#!/bin/bash
# round-2 ncu evidence: --set full pages of EVERY kernel family of the path (north star: "each kernel has a committed
# ncu capture reporting achieved HBM GB/s and tensor-pipe utilisation"), in situ (one bench step, fp16 mode, eager
# launches), + the launch list of the same step.  Summaries are produced offline by scripts/ncu_summary.py.
set -u
mkdir -p gpurun_out/ncu
N="ncu --set full --clock-control none --import-source on -f"
# main kernels, steady state (second generated frame onwards)
timeout -k 10 500 $N -k regex:gemm_tcgen05 --launch-skip 1170 -c 12 -o gpurun_out/ncu/r02_gemm_insitu python scripts/one_step.py 64 > gpurun_out/ncu/gemm.log 2>&1; echo "gemm rc=$?"
timeout -k 10 300 $N -k regex:spatial_attn --launch-skip 200 -c 2 -o gpurun_out/ncu/r02_spatial python scripts/one_step.py 64 > gpurun_out/ncu/spatial.log 2>&1; echo "spatial rc=$?"
timeout -k 10 300 $N -k regex:temporal_attn --launch-skip 200 -c 2 -o gpurun_out/ncu/r02_temporal python scripts/one_step.py 64 > gpurun_out/ncu/temporal.log 2>&1; echo "temporal rc=$?"
# small kernels of the generate path
timeout -k 10 300 $N -k regex:prep_kernel --launch-skip 400 -c 2 -o gpurun_out/ncu/r02_prep python scripts/one_step.py 64 > gpurun_out/ncu/prep.log 2>&1; echo "prep rc=$?"
timeout -k 10 300 $N -k "regex:embed_kernel|sample_kernel|remask_kernel|fill_i32" --launch-skip 6 -c 6 -o gpurun_out/ncu/r02_decode python scripts/one_step.py 64 > gpurun_out/ncu/decode.log 2>&1; echo "decode rc=$?"
# evaluate path: CE / count / logits transpose
timeout -k 10 300 $N -k "regex:ce_kernel|count_equal|check_masked|logits_transpose|relevant_weight" --launch-skip 4 -c 6 -o gpurun_out/ncu/r02_eval python scripts/bench_eval.py 8 > gpurun_out/ncu/eval.log 2>&1; echo "eval rc=$?"
# MAGVIT2 kernels (8 images: one workspace pass)
timeout -k 10 400 $N -k "regex:stem_conv|gn_partial|gn_finalize|gn_apply|depth_to_space|vq_head|vq_tail|out_conv" --launch-skip 150 -c 40 -o gpurun_out/ncu/r02_vq python scripts/bench_magvit.py 8 > gpurun_out/ncu/vq.log 2>&1; echo "vq rc=$?"
# launch list of one step
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/ncu/r02_launches.csv python scripts/one_step.py 64 > gpurun_out/ncu/launches.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/ncu/r02_launches.csv
ls -la gpurun_out/ncu
