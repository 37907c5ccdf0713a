#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary6.txt
cd tests
timeout -k 10 600 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/r6_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary6.txt
timeout -k 10 900 python -m pytest -q -s -m gpu test_gpu_model.py > ../gpurun_out/r6_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary6.txt
cd ..
timeout -k 10 300 python scripts/gemm_microbench.py > gpurun_out/gemm_micro_r6.jsonl 2> gpurun_out/gemm_micro.err; echo "micro rc=$?" >> gpurun_out/summary6.txt
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/bench_r6.json 2>> gpurun_out/bench_r6.err; echo "bench rc=$?" >> gpurun_out/summary6.txt
cat gpurun_out/summary6.txt
