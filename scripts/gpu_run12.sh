#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary12.txt
cd tests
timeout -k 10 600 python -m pytest -q -s -m gpu test_gpu_model.py -k "wide_model" > ../gpurun_out/r12_wide.log 2>&1; echo "wide rc=$?" >> ../gpurun_out/summary12.txt
cd ..
timeout -k 10 300 python scripts/bench_magvit.py 64 > gpurun_out/bench_magvit_r1.json 2> gpurun_out/bench_magvit.err; echo "magvit bench rc=$?" >> gpurun_out/summary12.txt
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 12 -c 12 -o gpurun_out/prof_r12_gemm python bench.py --steps 1 --warmup 0 --batch 14 --no-cpu-baseline --no-secondary > gpurun_out/ncu_r12_gemm.log 2>&1; echo "ncu gemm rc=$?" >> gpurun_out/summary12.txt
cat gpurun_out/summary12.txt; grep -E "rel|passed|failed" gpurun_out/r12_wide.log; cat gpurun_out/bench_magvit_r1.json; tail -3 gpurun_out/bench_magvit.err
