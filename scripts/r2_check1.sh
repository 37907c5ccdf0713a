#!/bin/bash
# round-2 first GPU validation: fp16 mode parity table + the existing GPU suite
set -u
mkdir -p gpurun_out
timeout -k 10 600 python scripts/parity_report.py --skip-long --modes fp16 bf16 tf32 > gpurun_out/r2_parity1.jsonl 2> gpurun_out/r2_parity1.err; echo "parity rc=$?"
cd tests && timeout -k 10 900 python -m pytest -q -x -m gpu . > ../gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?"; cd ..
tail -3 gpurun_out/r2_tests1.log
cat gpurun_out/r2_parity1.jsonl
tail -5 gpurun_out/r2_parity1.err
