#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary5.txt
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --chunk-tokens 32768 > gpurun_out/bench_r5_cat.json 2>> gpurun_out/bench_r5.err; echo "bench rc=$?" >> gpurun_out/summary5.txt
timeout -k 10 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --chunk-tokens 32768 --mode dense > gpurun_out/bench_r5_dense.json 2>> gpurun_out/bench_r5.err; echo "bench dense rc=$?" >> gpurun_out/summary5.txt
for nm in "proj+res(dual)" "fc1+gelu" "qkv"; do
  tag=$(echo $nm | tr -d '+()')
  timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 1 -o gpurun_out/prof_r5_$tag python scripts/gemm_microbench.py "$nm" 16384 > gpurun_out/ncu_r5_$tag.log 2>&1; echo "ncu $tag rc=$?" >> gpurun_out/summary5.txt
done
cat gpurun_out/summary5.txt
