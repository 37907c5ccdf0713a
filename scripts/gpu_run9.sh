#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary9.txt
cd tests
timeout -k 10 600 python -m pytest -q -s -m gpu test_gpu_magvit.py > ../gpurun_out/r9_magvit.log 2>&1; echo "magvit rc=$?" >> ../gpurun_out/summary9.txt
cd ..
for b in 4 16; do
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --batch $b > gpurun_out/bench_r9_b$b.json 2>> gpurun_out/bench_r9.err; echo "bench b$b rc=$?" >> gpurun_out/summary9.txt
done
cat gpurun_out/summary9.txt; grep -E "latents rel|decode rel|passed|failed" gpurun_out/r9_magvit.log
python - <<'PY'
import json
for b in (4,16):
    d=json.load(open(f'gpurun_out/bench_r9_b{b}.json'))
    print(b, round(d['value'],1),'frames/s', 'ms/step', round(d['ms_per_step'],1), 'host enqueue', round(d['host_enqueue_ms_per_step'],1), 'launches/step', d['gpu_launches']/3)
PY
