set -u
mkdir -p gpurun_out
cd tests; timeout 300 python -m pytest -q -x -m gpu test_gpu_kernels.py -k "linear" 2>&1 | tail -3; cd ..
B="--no-cpu-baseline --no-secondary"
run() { # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py $B > gpurun_out/e21_bench_$name.json 2> gpurun_out/e21_bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e21_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("$name", "ERR", e)
PY
}
run a X=1
run b X=1
timeout 200 python scripts/gemm_microbench.py "proj+res(dual)" 32768
timeout 200 python scripts/gemm_microbench.py "proj+res" 32768
timeout 200 python scripts/gemm_microbench.py "fc2+res" 32768
