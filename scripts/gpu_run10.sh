#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary10.txt
cd tests
timeout -k 10 600 python -m pytest -q -s -m gpu test_gpu_eval_driver.py > ../gpurun_out/r10_eval.log 2>&1; echo "eval rc=$?" >> ../gpurun_out/summary10.txt
cd ..
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:temporal_attn -s 1000 -c 2 -o gpurun_out/prof_r10_temporal $B > gpurun_out/ncu_r10_temporal.log 2>&1; echo "ncu temporal rc=$?" >> gpurun_out/summary10.txt
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:spatial_attn -s 1000 -c 1 -o gpurun_out/prof_r10_spatial $B > gpurun_out/ncu_r10_spatial.log 2>&1; echo "ncu spatial rc=$?" >> gpurun_out/summary10.txt
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6000 -c 6 -o gpurun_out/prof_r10_gemm $B > gpurun_out/ncu_r10_gemm.log 2>&1; echo "ncu gemm rc=$?" >> gpurun_out/summary10.txt
cat gpurun_out/summary10.txt; tail -5 gpurun_out/r10_eval.log
