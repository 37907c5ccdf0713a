#!/bin/bash
# One-shot single-GPU validation (run through gpurun): full GPU test suite, smoke, parity table of every precision mode,
# memcheck of the small-shape tests, the benchmark (both arms) and the auxiliary per-config throughput scripts.
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 1500 python -m pytest -q -x -rP -m gpu . > ../gpurun_out/check_tests.log 2>&1; echo "tests rc=$?" > ../gpurun_out/check_summary.txt; cd ..
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/check_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 900 python scripts/parity_report.py --out gpurun_out/check_parity.jsonl > /dev/null 2> gpurun_out/check_parity.err; echo "parity rc=$?" >> gpurun_out/check_summary.txt
cd tests && timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest -q -x -m gpu test_gpu_model.py -k "tiny_maskgit_tokens or fused_readout_sample and genie35m or head_dim_32_with" > ../gpurun_out/check_sanitizer.log 2>&1; echo "memcheck rc=$?" >> ../gpurun_out/check_summary.txt; cd ..
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/check_bench_reference.json 2> gpurun_out/check_bench.err; echo "reference arm rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 900 python bench.py > gpurun_out/check_bench.json 2>> gpurun_out/check_bench.err; echo "bench rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 300 python scripts/bench_magvit.py 64 > gpurun_out/check_bench_magvit.json 2>> gpurun_out/check_bench.err; echo "magvit rc=$?" >> gpurun_out/check_summary.txt
timeout -k 10 300 python scripts/bench_generate.py --layers 32 --d-model 256 --heads 8 --batch 64 --maskgit-steps 2 > gpurun_out/check_bench_35m.json 2>> gpurun_out/check_bench.err; echo "35m rc=$?" >> gpurun_out/check_summary.txt
cat gpurun_out/check_summary.txt; tail -3 gpurun_out/check_tests.log
