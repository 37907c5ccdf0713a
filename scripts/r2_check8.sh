#!/bin/bash
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 1500 python -m pytest -q -x -rP -m gpu . > ../gpurun_out/r2_tests8.log 2>&1; echo "tests rc=$?"; cd ..
tail -3 gpurun_out/r2_tests8.log; grep -h "hd32 qk_norm" gpurun_out/r2_tests8.log
bash scripts/r2_ncu.sh
