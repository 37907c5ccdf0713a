#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary4.txt
cd tests
timeout -k 10 600 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/r4_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary4.txt
timeout -k 10 900 python -m pytest -q -s -m gpu test_gpu_model.py > ../gpurun_out/r4_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary4.txt
cd ..
timeout -k 10 300 python scripts/gemm_microbench.py > gpurun_out/gemm_micro_r4.jsonl 2> gpurun_out/gemm_micro.err; echo "micro rc=$?" >> gpurun_out/summary4.txt
for pdl in 1 0; do for ct in 16384 32768 65536; do
  GENIE_B200_PDL=$pdl timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --chunk-tokens $ct > gpurun_out/bench_r4_pdl${pdl}_$ct.json 2>> gpurun_out/bench_r4.err; echo "pdl $pdl chunk $ct rc=$?" >> gpurun_out/summary4.txt
done; done
cat gpurun_out/summary4.txt
