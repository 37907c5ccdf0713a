set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --no-secondary"
run() { # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py $B > gpurun_out/e19_bench_$name.json 2> gpurun_out/e19_bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e19_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("$name", "ERR", e)
PY
}
run s0 GENIE_B200_STREAM_HINT=0
run s1 GENIE_B200_STREAM_HINT=1
run s0b GENIE_B200_STREAM_HINT=0
run s1b GENIE_B200_STREAM_HINT=1
cd tests; timeout 300 python -m pytest -q -x -m gpu test_gpu_kernels.py -k "linear or spatial" 2>&1 | tail -3
