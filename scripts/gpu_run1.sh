#!/bin/bash
# first GPU bring-up: each stage under its own timeout so a hung kernel cannot eat the whole call
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
cd tests
timeout -k 10 300 python -m pytest -q -m gpu test_gpu_kernels.py -k "simt or sample or remask or cross_entropy" > ../gpurun_out/t1_safe.log 2>&1; echo "safe rc=$?" >> ../gpurun_out/summary.txt
timeout -k 10 300 python -m pytest -q -m gpu test_gpu_kernels.py -k "bf16_store" > ../gpurun_out/t2_tc_store.log 2>&1; echo "tc_store rc=$?" >> ../gpurun_out/summary.txt
timeout -k 10 300 python -m pytest -q -m gpu test_gpu_kernels.py -k "gelu_and_residual or tf32 or bitwise" > ../gpurun_out/t3_tc_rest.log 2>&1; echo "tc_rest rc=$?" >> ../gpurun_out/summary.txt
timeout -k 10 900 python -m pytest -q -s -m gpu test_gpu_model.py > ../gpurun_out/t4_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary.txt
cd ..
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/t5_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t1_safe.log gpurun_out/t2_tc_store.log gpurun_out/t3_tc_rest.log gpurun_out/t4_model.log gpurun_out/t5_smoke.log
