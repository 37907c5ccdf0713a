#!/bin/bash
# Round-1 experiment batch 1: temporal attention v2 + 256-wide dual residual epilogue (fold_ln) A/B.
set -u
mkdir -p gpurun_out
S=gpurun_out/e1_summary.txt; : > $S
cd tests
timeout -k 10 500 python -m pytest -q -x -m gpu test_gpu_kernels.py -k "linear" > ../gpurun_out/e1_kernels.log 2>&1; echo "kernels rc=$?" >> ../$S
timeout -k 10 700 python -m pytest -q -x -s -m gpu test_gpu_model.py -k "temporal_v2 or folded or production_maskgit or wide_model or cuda_graph" > ../gpurun_out/e1_model.log 2>&1; echo "model rc=$?" >> ../$S
cd ..
B="--no-cpu-baseline --no-secondary"
timeout -k 10 300 python bench.py $B > gpurun_out/e1_bench_v2.json 2> gpurun_out/e1_bench_v2.err; echo "bench v2 rc=$?" >> $S
GENIE_B200_TEMPORAL_V2=0 timeout -k 10 300 python bench.py $B > gpurun_out/e1_bench_legacy.json 2> gpurun_out/e1_bench_legacy.err; echo "bench legacy rc=$?" >> $S
timeout -k 10 300 python bench.py $B --fold-ln > gpurun_out/e1_bench_fold.json 2> gpurun_out/e1_bench_fold.err; echo "bench fold rc=$?" >> $S
timeout -k 10 300 python bench.py $B --mode dense --steps 2 > gpurun_out/e1_bench_dense.json 2> gpurun_out/e1_bench_dense.err; echo "bench dense rc=$?" >> $S
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:temporal_attn_v2 --launch-skip 500 -c 2 -f -o gpurun_out/e1_temporal_v2 python bench.py --steps 1 --warmup 1 $B --no-graphs > gpurun_out/e1_ncu_temporal.log 2>&1; echo "ncu temporal rc=$?" >> $S
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 700 --csv --log-file gpurun_out/e1_launches.csv python bench.py --steps 1 --warmup 1 $B --no-graphs > gpurun_out/e1_ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $S
cd tests
timeout -k 10 1200 python -m pytest -q -x -m gpu . > ../gpurun_out/e1_tests_full.log 2>&1; echo "full tests rc=$?" >> ../$S
cd ..
cat $S; tail -3 gpurun_out/e1_tests_full.log
for f in v2 legacy fold dense; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e1_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()})
except Exception as e: print("$f", "ERR", e)
PY
done
