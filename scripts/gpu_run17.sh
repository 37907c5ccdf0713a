#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary17.txt
cd tests
timeout -k 10 900 python -m pytest -q -s -m gpu test_gpu_model.py > ../gpurun_out/r17_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary17.txt
cd ..
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/bench_r17.json 2>> gpurun_out/bench_r17.err; echo "bench rc=$?" >> gpurun_out/summary17.txt
timeout -k 10 300 python scripts/bench_eval.py 32 > gpurun_out/bench_eval_r1.json 2>> gpurun_out/bench_r17.err; echo "eval bench rc=$?" >> gpurun_out/summary17.txt
cat gpurun_out/summary17.txt; grep -E "passed|failed|teacher-forced CE|Error" gpurun_out/r17_model.log | head; cat gpurun_out/bench_eval_r1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r17.json'))
print(round(d['value'],1),'frames/s', 'ms/step', round(d['ms_per_step'],1), 'gemm TF', round(d['roofline']['achieved'],1), 'e2e', round(d['e2e']['value'],1))
print('   ', {k: round(v['ms_per_step'],1) for k,v in d['roofline']['kernel_ms_by_category'].items()})
PY
