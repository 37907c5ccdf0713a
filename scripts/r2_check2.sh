#!/bin/bash
# round-2 GPU validation 2: full GPU suite, smoke, parity report (all modes, long fixtures), default bench
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 1200 python -m pytest -q -x -m gpu . > ../gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?"; cd ..
tail -5 gpurun_out/r2_tests2.log
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke2.log
timeout -k 10 900 python scripts/parity_report.py --out gpurun_out/r2_parity2.jsonl > /dev/null 2> gpurun_out/r2_parity2.err; echo "parity rc=$?"
cat gpurun_out/r2_parity2.jsonl
timeout -k 10 900 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
cat gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
