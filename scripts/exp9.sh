set -u
mkdir -p gpurun_out
cd tests
timeout -k 5 150 python -m pytest -q -x -m gpu test_gpu_kernels.py -k spatial -s 2>&1 | tail -25 > ../gpurun_out/e9_kernel.log
rc=$?
cd ..
cat gpurun_out/e9_kernel.log
if grep -q "passed" gpurun_out/e9_kernel.log && ! grep -q "failed" gpurun_out/e9_kernel.log; then
  cd tests; timeout -k 10 900 python -m pytest -q -x -m gpu test_gpu_model.py 2>&1 | tail -8 > ../gpurun_out/e9_model.log; cd ..
  cat gpurun_out/e9_model.log
  for tc in 1 0; do
    GENIE_B200_SPATIAL_TC=$tc timeout -k 10 300 python bench.py --lanes 1 --no-cpu-baseline --no-secondary > gpurun_out/e9_bench_tc$tc.json 2> gpurun_out/e9_bench_tc$tc.err
    echo "bench tc=$tc rc=$?"
  done
  python - <<'PY'
import json
for l in (1,0):
    try:
        d=json.loads(open(f"gpurun_out/e9_bench_tc{l}.json").read().strip().splitlines()[-1])
        print(l, round(d["value"],1), round(d["ms_per_step"],2), {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"])
    except Exception as e: print(l, "ERR", e)
PY
fi
