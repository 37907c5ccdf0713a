set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --no-secondary"
timeout -k 10 300 python bench.py $B > gpurun_out/e15_bench_default.json 2> gpurun_out/e15_bench_default.err; echo "default rc=$?"
timeout -k 10 300 python bench.py $B --fold-ln > gpurun_out/e15_bench_fold.json 2> gpurun_out/e15_bench_fold.err; echo "fold rc=$?"
GENIE_B200_PAIR_PROJ=1 timeout -k 10 300 python bench.py $B > gpurun_out/e15_bench_pairproj.json 2> gpurun_out/e15_bench_pairproj.err; echo "pairproj rc=$?"
for f in default fold pairproj; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e15_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("$f", "ERR", e)
PY
done
timeout -k 10 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/e15_launches.csv python scripts/one_step.py > gpurun_out/e15_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
tail -2 gpurun_out/e15_ncu_launches.log
wc -l gpurun_out/e15_launches.csv
