#!/bin/bash
# One-shot GPU validation used during development (run through gpurun):
#   full GPU test suite, smoke, the benchmark (both arms) and the auxiliary throughput scripts.
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 1500 python -m pytest -q -x -m gpu . > ../gpurun_out/check_tests.log 2>&1; echo "tests rc=$?" > ../gpurun_out/check_summary.txt; cd ..
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/check_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 900 python bench.py > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 600 python bench.py --impl reference > gpurun_out/check_bench_reference.json 2>> gpurun_out/check_bench.err; echo "reference arm rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 300 python scripts/bench_magvit.py 64 > gpurun_out/check_bench_magvit.json 2>> gpurun_out/check_bench.err; echo "magvit rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 300 python scripts/bench_eval.py 32 > gpurun_out/check_bench_eval.json 2>> gpurun_out/check_bench.err; echo "eval rc=$?" >> gpurun_out/check_summary.txt
cat gpurun_out/check_summary.txt; tail -3 gpurun_out/check_tests.log
