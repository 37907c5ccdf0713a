set -u
mkdir -p gpurun_out
cd tests; timeout 600 python -m pytest -q -x -m gpu test_gpu_model.py -k "lanes or graph" 2>&1 | tail -5 > ../gpurun_out/e8_tests.log; cd ..
for l in 1 2 3; do
  timeout -k 10 300 python bench.py --lanes $l --no-cpu-baseline --no-secondary > gpurun_out/e8_bench_l$l.json 2> gpurun_out/e8_bench_l$l.err
  echo "lanes $l rc=$?" >> gpurun_out/e8_tests.log
done
cat gpurun_out/e8_tests.log
python - <<'PY'
import json
for l in (1,2,3):
    try:
        d=json.loads(open(f"gpurun_out/e8_bench_l{l}.json").read().strip().splitlines()[-1])
        print(l, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"])
    except Exception as e: print(l, "ERR", e)
PY
