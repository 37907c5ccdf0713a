#!/bin/bash
# Experiment 3: where does the linear kernel's time go?  (epilogue / loads / MMA issue ablations; garbage results)
set -u
mkdir -p gpurun_out
: > gpurun_out/e3_micro.jsonl
for pair in 1 0; do for dbg in 0 1 2 3; do
  for shape in qkv fc1+gelu; do
    for M in 32768 262144; do
      echo "{\"pair\": $pair, \"dbg\": $dbg}" >> gpurun_out/e3_micro.jsonl
      GENIE_B200_PAIR=$pair GENIE_B200_GEMM_DEBUG=$dbg timeout -k 5 120 python scripts/gemm_microbench.py "$shape" $M >> gpurun_out/e3_micro.jsonl 2>> gpurun_out/e3_micro.err
    done
  done
done; done
cat gpurun_out/e3_micro.jsonl
