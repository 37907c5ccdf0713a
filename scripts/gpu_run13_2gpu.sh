#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary13.txt
nvidia-smi -L > gpurun_out/r13_gpus.txt
cd tests
timeout -k 10 600 python -m pytest -q -s -m gpu test_gpu_eval_driver.py > ../gpurun_out/r13_eval2.log 2>&1; echo "eval2 rc=$?" >> ../gpurun_out/summary13.txt
timeout -k 10 300 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/r13_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary13.txt
cd ..
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r13_2gpu.json 2> gpurun_out/bench_r13.err; echo "bench2 rc=$?" >> gpurun_out/summary13.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r13_ref.json 2>> gpurun_out/bench_r13.err; echo "ref2 rc=$?" >> gpurun_out/summary13.txt
cat gpurun_out/summary13.txt; tail -3 gpurun_out/r13_eval2.log; tail -2 gpurun_out/r13_kernels.log; cat gpurun_out/bench_r13_2gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value'],1), 'frames/s', round(d['ms_per_step'],1),'ms/step e2e', round(d['e2e']['value'],1))"; cat gpurun_out/bench_r13_ref.json | cut -c1-400; tail -3 gpurun_out/bench_r13.err
