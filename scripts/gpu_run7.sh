#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary7.txt
cd tests
timeout -k 10 600 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/r7_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary7.txt
timeout -k 10 900 python -m pytest -q -s -m gpu test_gpu_model.py > ../gpurun_out/r7_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary7.txt
cd ..
timeout -k 10 300 python scripts/gemm_microbench.py "fc1+gelu" > gpurun_out/gemm_micro_r7.jsonl 2> gpurun_out/gemm_micro.err; echo "micro rc=$?" >> gpurun_out/summary7.txt
for ct in 8192 16384 32768; do
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --chunk-tokens $ct > gpurun_out/bench_r7_$ct.json 2>> gpurun_out/bench_r7.err; echo "bench $ct rc=$?" >> gpurun_out/summary7.txt
done
cat gpurun_out/summary7.txt
