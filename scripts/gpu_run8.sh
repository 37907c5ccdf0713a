#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary8.txt
cd tests
timeout -k 10 600 python -m pytest -q -s -m gpu test_gpu_magvit.py > ../gpurun_out/r8_magvit.log 2>&1; echo "magvit rc=$?" >> ../gpurun_out/summary8.txt
timeout -k 10 600 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/r8_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary8.txt
timeout -k 10 900 python -m pytest -q -m gpu test_gpu_model.py > ../gpurun_out/r8_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary8.txt
cd ..
timeout -k 10 300 python scripts/gemm_microbench.py "fc1+gelu" > gpurun_out/gemm_micro_r8.jsonl 2> gpurun_out/gemm_micro.err; echo "micro rc=$?" >> gpurun_out/summary8.txt
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/bench_r8.json 2>> gpurun_out/bench_r8.err; echo "bench rc=$?" >> gpurun_out/summary8.txt
cat gpurun_out/summary8.txt; tail -30 gpurun_out/r8_magvit.log
