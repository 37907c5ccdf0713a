#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary18.txt
timeout -k 10 900 python bench.py > gpurun_out/bench_final_r1.json 2> gpurun_out/bench_final.err; echo "bench rc=$?" >> gpurun_out/summary18.txt
timeout -k 10 900 python bench.py --impl reference > gpurun_out/bench_final_ref_r1.json 2>> gpurun_out/bench_final.err; echo "ref rc=$?" >> gpurun_out/summary18.txt
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary18.txt
cat gpurun_out/summary18.txt; cat gpurun_out/bench_final_r1.json; cat gpurun_out/bench_final_ref_r1.json | cut -c1-300; tail -2 gpurun_out/smoke_final.log
