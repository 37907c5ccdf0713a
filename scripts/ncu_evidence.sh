#!/bin/bash
# round-2 ncu evidence: --set full of EVERY kernel family of the path (north star: "each kernel has a committed ncu
# capture reporting achieved HBM GB/s and tensor-pipe utilisation"), in situ (one bench step, default fp16 mode, eager
# launches), + the launch list of the same step.  The raw pages are exported to CSV on the box (the .ncu-rep files are
# too large to travel); scripts/ncu_summary.py turns them into profiles/r02_ncu_kernels.{md,json}.
set -u
D=gpurun_out/ncu
mkdir -p $D
N="ncu --set full --clock-control none -f"
cap() {   # name, kernel regex, launch-skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout -k 10 600 $N -k "regex:$rx" --launch-skip $skip -c $cnt -o $D/$name "$@" > $D/$name.log 2>&1
  echo "$name rc=$?"
  ncu -i $D/$name.ncu-rep --page raw --csv > $D/${name}_raw.csv 2>/dev/null
  rm -f $D/$name.ncu-rep
  gzip -f $D/${name}_raw.csv
}
cap r02_gemm_insitu gemm_tcgen05 1170 12 python scripts/one_step.py 64
cap r02_spatial spatial_attn 200 2 python scripts/one_step.py 64
cap r02_temporal temporal_attn 200 2 python scripts/one_step.py 64
cap r02_prep prep_kernel 400 2 python scripts/one_step.py 64
cap r02_decode "embed_kernel|readout_sample|remask_kernel|fill_i32|sample_kernel" 4 8 python scripts/one_step.py 64
cap r02_eval "ce_kernel|count_equal|check_masked|logits_transpose|relevant_weight|readout_sample" 4 8 python scripts/bench_eval.py 8
cap r02_vq "stem_conv|gn_partial|gn_apply|depth_to_space|vq_head|vq_tail|out_conv" 150 28 python scripts/bench_magvit.py 8
cap r02_vq_conv gemm_tcgen05 150 6 python scripts/bench_magvit.py 8
cap r02_vq_extra "depth_to_space|vq_tail|out_conv" 4 6 python scripts/bench_magvit.py 8
# kernels the bench step / evaluate script do not launch (maskgit_generate's precondition check and logits layout,
# forward's relevant-position weights, the temperature > 0 sampler)
cap r02_extra "check_masked|logits_transpose|relevant_weight|sample_kernel|fill_i32|::ce_kernel" 6 10 python scripts/ncu_extra_driver.py
timeout -k 10 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $D/r02_launches.csv python scripts/one_step.py 64 > $D/launches.log 2>&1; echo "launch list rc=$?"
gzip -f $D/r02_launches.csv
ls -la $D; du -sh gpurun_out
