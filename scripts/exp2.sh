#!/bin/bash
# Experiment batch 2: CTA-pair (cta_group::2) GEMM tiles A/B.
set -u
mkdir -p gpurun_out
S=gpurun_out/e2_summary.txt; : > $S
cd tests
timeout -k 10 400 python -m pytest -q -x -m gpu test_gpu_kernels.py > ../gpurun_out/e2_kernels.log 2>&1; echo "kernels rc=$?" >> ../$S
cd ..
if grep -q "rc=0" $S; then
  for M in 16384 32768 262144; do
    timeout -k 10 200 python scripts/gemm_microbench.py "" $M >> gpurun_out/e2_micro_pair.jsonl 2>> gpurun_out/e2_micro.err
    GENIE_B200_PAIR=0 timeout -k 10 200 python scripts/gemm_microbench.py "" $M >> gpurun_out/e2_micro_single.jsonl 2>> gpurun_out/e2_micro.err
  done
  echo "micro done" >> $S
  B="--no-cpu-baseline --no-secondary"
  timeout -k 10 300 python bench.py $B > gpurun_out/e2_bench_pair.json 2> gpurun_out/e2_bench_pair.err; echo "bench pair rc=$?" >> $S
  GENIE_B200_PAIR=0 timeout -k 10 300 python bench.py $B > gpurun_out/e2_bench_single.json 2> gpurun_out/e2_bench_single.err; echo "bench single rc=$?" >> $S
  timeout -k 10 300 python bench.py $B --mode dense --steps 2 > gpurun_out/e2_bench_dense.json 2> gpurun_out/e2_bench_dense.err; echo "bench dense rc=$?" >> $S
  cd tests
  timeout -k 10 900 python -m pytest -q -x -m gpu . > ../gpurun_out/e2_tests_full.log 2>&1; echo "full tests rc=$?" >> ../$S
  cd ..
  timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 700 -c 6 -f -o gpurun_out/e2_gemm_pair python bench.py --steps 1 --warmup 1 $B --no-graphs > gpurun_out/e2_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" >> $S
fi
cat $S; tail -5 gpurun_out/e2_kernels.log; tail -3 gpurun_out/e2_tests_full.log 2>/dev/null
for f in pair single dense; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["roofline"]["achieved"], d["clocks"])
except Exception as e: print("$f", "ERR", e)
PY
done
python - <<'PY'
import json
for f in ("pair","single"):
    try:
        for l in open(f"gpurun_out/e2_micro_{f}.jsonl"):
            d=json.loads(l)
            if not d["l2_flush"]: print(f, d["M"], d["name"], d["us"], d["tflops"])
    except Exception as e: print(f, "ERR", e)
PY
