set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --no-secondary"
for p in 16 32 48; do
GENIE_B200_L2_PERSIST=$p timeout -k 10 300 python bench.py $B > gpurun_out/e17_bench_p$p.json 2> gpurun_out/e17_bench_p$p.err; echo "persist=$p rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e17_bench_p$p.json").read().strip().splitlines()[-1])
    print("p$p", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("p$p", "ERR", e)
PY
done
