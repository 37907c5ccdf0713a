#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python scripts/gemm_microbench.py > gpurun_out/gemm_micro_r1.jsonl 2> gpurun_out/gemm_micro.err; echo "micro rc=$?" >> gpurun_out/summary3.txt
for ct in 8192 16384 32768 65536; do
  timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --chunk-tokens $ct > gpurun_out/bench_chunk_$ct.json 2>> gpurun_out/bench_chunk.err; echo "chunk $ct rc=$?" >> gpurun_out/summary3.txt
done
cat gpurun_out/summary3.txt
