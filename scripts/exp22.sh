set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --no-secondary"
run() { # name, extra bench args, env...
  name=$1; shift; extra=$1; shift
  env "$@" timeout -k 10 300 python bench.py $B $extra > gpurun_out/e22_bench_$name.json 2> gpurun_out/e22_bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e22_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("$name", "ERR", e)
PY
}
run base "" X=1
run proj256 "" GENIE_B200_PROJ256=1
run chunk16k "--chunk-tokens 16384" X=1
run p40 "" GENIE_B200_L2_PERSIST=40
run p56 "" GENIE_B200_L2_PERSIST=56
run base2 "" X=1
