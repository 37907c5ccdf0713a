#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary16.txt
for ct in 16384 24576 49152; do
timeout -k 10 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --chunk-tokens $ct > gpurun_out/bench_r16_$ct.json 2>> gpurun_out/bench_r16.err; echo "bench $ct rc=$?" >> gpurun_out/summary16.txt
done
cd tests; timeout -k 10 600 python -m pytest -q -s -m gpu test_gpu_model.py -k "folded or tiny_logits" > ../gpurun_out/r16_model.log 2>&1; echo "model rc=$?" >> ../gpurun_out/summary16.txt; cd ..
cat gpurun_out/summary16.txt; grep -E "fold vs|passed|failed" gpurun_out/r16_model.log
python - <<'PY'
import json
for ct in (16384,24576,49152):
    d=json.load(open(f'gpurun_out/bench_r16_{ct}.json'))
    print(ct, round(d['value'],1),'frames/s', 'ms/step', round(d['ms_per_step'],1), 'gemm TF', round(d['roofline']['achieved'],1))
    print('   ', {k: round(v['ms_per_step'],1) for k,v in d['roofline']['kernel_ms_by_category'].items()})
PY
