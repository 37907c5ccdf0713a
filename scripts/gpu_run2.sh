#!/bin/bash
mkdir -p gpurun_out
cd tests
timeout -k 10 600 python -m pytest -q -s -m gpu test_gpu_kernels.py -k "cross_entropy or tf32" > ../gpurun_out/r2_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary2.txt
timeout -k 10 900 python -m pytest -q -s -m gpu test_gpu_model.py -k "tf32" > ../gpurun_out/r2_model_tf32.log 2>&1; echo "model_tf32 rc=$?" >> ../gpurun_out/summary2.txt
cd ..
timeout -k 10 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?" >> gpurun_out/summary2.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-secondary > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/summary2.txt
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 400 -c 4 -o gpurun_out/prof_gemm_r1 python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt; cat gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
