set -u
mkdir -p gpurun_out
cd tests
timeout -k 5 150 python -m pytest -q -x -m gpu test_gpu_kernels.py -k spatial -s 2>&1 | tail -25 > ../gpurun_out/e10_kernel.log
cd ..
cat gpurun_out/e10_kernel.log
if grep -q "passed" gpurun_out/e10_kernel.log && ! grep -q "failed" gpurun_out/e10_kernel.log; then
  timeout -k 10 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/e10_bench.json 2> gpurun_out/e10_bench.err
  echo "bench rc=$?"
  python - <<'PY'
import json
d=json.loads(open("gpurun_out/e10_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"])
PY
  cd tests; timeout -k 10 900 python -m pytest -q -x -m gpu test_gpu_model.py 2>&1 | tail -4; cd ..
fi
