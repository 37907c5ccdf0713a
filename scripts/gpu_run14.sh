#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary14.txt
cd tests
timeout -k 10 300 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/r14_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary14.txt
timeout -k 10 900 python -m pytest -q -s -m gpu test_gpu_model.py > ../gpurun_out/r14_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary14.txt
cd ..
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/bench_r14_fold.json 2>> gpurun_out/bench_r14.err; echo "bench fold rc=$?" >> gpurun_out/summary14.txt
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-fold-ln > gpurun_out/bench_r14_nofold.json 2>> gpurun_out/bench_r14.err; echo "bench nofold rc=$?" >> gpurun_out/summary14.txt
cat gpurun_out/summary14.txt; tail -2 gpurun_out/r14_kernels.log; grep -E "bf16|passed|failed|Error" gpurun_out/r14_model.log | head -30
python - <<'PY'
import json
for f in ('fold','nofold'):
    d=json.load(open(f'gpurun_out/bench_r14_{f}.json'))
    print(f, round(d['value'],1),'frames/s', 'ms/step', round(d['ms_per_step'],1), 'gemm TF', round(d['roofline']['achieved'],1), 'e2e', round(d['e2e']['value'],1))
    for k,v in d['roofline']['kernel_ms_by_category'].items():
        print(f"   {k:14s} {v['ms_per_step']:8.2f} ms  {v['launches_per_step']:7.0f} launches  avg {1e3*v['ms_per_step']/max(v['launches_per_step'],1):7.1f} us")
PY
