#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary15.txt
cd tests
timeout -k 10 300 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/r15_kernels.log 2>&1; echo "kernels rc=$?" >> ../gpurun_out/summary15.txt
cd ..
timeout -k 10 300 python scripts/gemm_microbench.py > gpurun_out/gemm_micro_r15.jsonl 2> gpurun_out/gemm_micro.err; echo "micro rc=$?" >> gpurun_out/summary15.txt
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-fold-ln > gpurun_out/bench_r15_nofold.json 2>> gpurun_out/bench_r15.err; echo "bench nofold rc=$?" >> gpurun_out/summary15.txt
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/bench_r15_fold.json 2>> gpurun_out/bench_r15.err; echo "bench fold rc=$?" >> gpurun_out/summary15.txt
cat gpurun_out/summary15.txt; tail -2 gpurun_out/r15_kernels.log
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/gemm_micro_r15.jsonl')]
old={ (r['M'],r['name'],r['l2_flush']):r for r in map(json.loads, open('gpurun_out/gemm_micro_r6.jsonl'))}
for r in rows:
    if r['l2_flush'] or r['M'] in (4096,): continue
    o=old.get((r['M'],r['name'],False),{'tflops':0})
    print(f"{r['M']:7d} {r['name']:16s} {r['us']:9.1f} {r['tflops']:7.1f} ({o['tflops']:7.1f})")
for f in ('nofold','fold'):
    d=json.load(open(f'gpurun_out/bench_r15_{f}.json'))
    print(f, round(d['value'],1),'frames/s', 'ms/step', round(d['ms_per_step'],1), 'gemm TF', round(d['roofline']['achieved'],1), 'e2e', round(d['e2e']['value'],1))
    for k,v in d['roofline']['kernel_ms_by_category'].items():
        print(f"   {k:14s} {v['ms_per_step']:8.2f} ms  {v['launches_per_step']:7.0f} launches  avg {1e3*v['ms_per_step']/max(v['launches_per_step'],1):7.1f} us")
PY
