#!/bin/bash
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 1500 python -m pytest -q -x -rP -m gpu . > ../gpurun_out/r2_tests7.log 2>&1; echo "tests rc=$?"; cd ..
tail -3 gpurun_out/r2_tests7.log; grep -h "temporal v2 vs\|magvit f\|magvit b" gpurun_out/r2_tests7.log
timeout -k 10 300 python scripts/bench_generate.py --layers 32 --d-model 256 --heads 8 --batch 64 --maskgit-steps 2 > gpurun_out/r2_bench_35m_b.json 2> gpurun_out/r2_bench_35m_b.err; echo "35m rc=$?"; cat gpurun_out/r2_bench_35m_b.json
GENIE_B200_TEMPORAL_V2=0 timeout -k 10 300 python scripts/bench_generate.py --layers 32 --d-model 256 --heads 8 --batch 64 --maskgit-steps 2 > gpurun_out/r2_bench_35m_tv1.json 2>> gpurun_out/r2_bench_35m_b.err; echo "35m tv1 rc=$?"; cat gpurun_out/r2_bench_35m_tv1.json
GENIE_PRECISION=fp16 timeout -k 10 300 python scripts/bench_magvit.py 64 > gpurun_out/r2_bench_magvit_fp16c.json 2> gpurun_out/r2_bench_magvit_fp16c.err; cat gpurun_out/r2_bench_magvit_fp16c.json
