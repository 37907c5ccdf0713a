set -u
mkdir -p gpurun_out
python - <<'PY'
import ctypes, torch
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_int()
for name, a in (("MaxPersistingL2CacheSize", 108), ("MaxAccessPolicyWindowSize", 109), ("L2CacheSize", 38)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0); print(name, v.value)
PY
B="--no-cpu-baseline --no-secondary"
for p in 1 0; do
GENIE_B200_L2_PERSIST=$p timeout -k 10 300 python bench.py $B > gpurun_out/e16_bench_p$p.json 2> gpurun_out/e16_bench_p$p.err; echo "persist=$p rc=$?"
done
for f in p1 p0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e16_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "frames/s", round(d["ms_per_step"],1), "ms", {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("$f", "ERR", e)
PY
done
cd tests; timeout 600 python -m pytest -q -x -m gpu test_gpu_model.py -k "graph or lanes or production" 2>&1 | tail -3
