set -u
mkdir -p gpurun_out
cd tests
timeout -k 5 200 python -m pytest -q -x -m gpu test_gpu_kernels.py -k spatial -s 2>&1 | tail -25 > ../gpurun_out/e14_kernel.log
cd ..
cat gpurun_out/e14_kernel.log
if grep -q "passed" gpurun_out/e14_kernel.log && ! grep -q "failed" gpurun_out/e14_kernel.log; then
  timeout 120 python scripts/spatial_microbench.py 128 8 20 0 > gpurun_out/e14_micro.jsonl 2>&1
  timeout 120 python scripts/spatial_microbench.py 64 8 20 0 >> gpurun_out/e14_micro.jsonl 2>&1
  GENIE_B200_SPATIAL_TC=1 timeout 120 python scripts/spatial_microbench.py 128 8 20 0 >> gpurun_out/e14_micro.jsonl 2>&1
  cat gpurun_out/e14_micro.jsonl
  timeout -k 10 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/e14_bench.json 2> gpurun_out/e14_bench.err
  echo "bench rc=$?"
  python - <<'PY'
import json
d=json.loads(open("gpurun_out/e14_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v["ms_per_step"],1) for k,v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"])
PY
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:spatial_attn_tc_persistent -c 2 -o gpurun_out/e14_spatial_tcp python scripts/spatial_microbench.py 128 8 1 0 > gpurun_out/e14_ncu.log 2>&1
  echo "ncu rc=$?"
fi
