#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary11.txt
cd tests
timeout -k 10 900 python -m pytest -q -m gpu test_gpu_model.py > ../gpurun_out/r11_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary11.txt
cd ..
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r11.json 2>> gpurun_out/bench_r11.err; echo "bench rc=$?" >> gpurun_out/summary11.txt
cat gpurun_out/summary11.txt; tail -3 gpurun_out/r11_model.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r11.json'))
print(round(d['value'],1),'frames/s', 'ms/step', round(d['ms_per_step'],1), 'profiled', round(d['roofline']['profiled_ms_per_step'],1), 'gemm TF', round(d['roofline']['achieved'],1), 'e2e', round(d['e2e']['value'],1), 'dense', d['secondary'])
for k,v in d['roofline']['kernel_ms_by_category'].items():
    print(f"   {k:14s} {v['ms_per_step']:8.2f} ms  {v['launches_per_step']:7.0f} launches  avg {1e3*v['ms_per_step']/max(v['launches_per_step'],1):7.1f} us")
PY
