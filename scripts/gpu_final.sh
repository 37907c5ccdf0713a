#!/bin/bash
# Round-end evidence run (through gpurun): in-situ ncu --set full capture of 12 consecutive tcgen05 GEMM launches of a
# decode forward at M = 32768 rows (second generated frame, MaskGIT step 0), then the full validation of gpu_check.sh.
set -u
mkdir -p gpurun_out
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 1170 -c 12 -f \
  -o gpurun_out/final_gemm_insitu python scripts/one_step.py 64 > gpurun_out/final_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ncu -i gpurun_out/final_gemm_insitu.ncu-rep --page raw --csv > gpurun_out/final_gemm_insitu_raw.csv 2>/dev/null
bash scripts/gpu_check.sh
