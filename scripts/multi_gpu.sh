#!/bin/bash
# multi-GPU validation (run with gpurun --gpus N): bench.py under torchrun (generate + evaluate leg with the NCCL
# all-reduce), the 2-rank tests, and the other BASELINE configs per GPU count.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  cd tests && timeout -k 10 900 python -m pytest -q -x -rP -m gpu test_gpu_cli.py test_gpu_eval_driver.py test_gpu_model.py -k "two_ranks or nccl or two_devices" > ../gpurun_out/multi_tests.log 2>&1; echo "multi tests rc=$?"; cd ..
  tail -3 gpurun_out/multi_tests.log
fi
timeout -k 10 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err; echo "bench N=$N rc=$?"
cat gpurun_out/multi_bench_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['evaluate'], d['e2e'])"
timeout -k 10 600 $TR scripts/bench_generate.py --layers 40 --d-model 1024 --heads 16 --batch 16 --maskgit-steps 8 > gpurun_out/multi_bench_700m_n$N.json 2> gpurun_out/multi_bench_700m_n$N.err; echo "700m rc=$?"; cat gpurun_out/multi_bench_700m_n$N.json
timeout -k 10 600 $TR scripts/bench_magvit.py 64 > gpurun_out/multi_bench_magvit_n$N.json 2> gpurun_out/multi_bench_magvit_n$N.err; echo "magvit rc=$?"; cat gpurun_out/multi_bench_magvit_n$N.json
timeout -k 10 600 $TR scripts/bench_eval.py 32 > gpurun_out/multi_bench_eval_n$N.json 2> gpurun_out/multi_bench_eval_n$N.err; echo "eval rc=$?"; cat gpurun_out/multi_bench_eval_n$N.json
tail -3 gpurun_out/multi_bench_n$N.err
