#!/bin/bash
# Final-tree validation of the second round-2 session (one gpurun call): full GPU suite, smoke, both bench arms, MAGVIT2.
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 600 python -m pytest -q -x -m gpu . > ../gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" > ../gpurun_out/f_summary.txt; cd ..
timeout -k 10 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_summary.txt
timeout -k 10 120 python scripts/bench_magvit.py 64 > gpurun_out/f_bench_magvit.json 2> gpurun_out/f_bench.err; echo "magvit rc=$?" >> gpurun_out/f_summary.txt
timeout -k 10 400 python bench.py > gpurun_out/f_bench.json 2>> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_summary.txt
timeout -k 10 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2>> gpurun_out/f_bench.err; echo "reference arm rc=$?" >> gpurun_out/f_summary.txt
cat gpurun_out/f_summary.txt; tail -3 gpurun_out/f_tests.log; tail -2 gpurun_out/f_smoke.log
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"]["logits_rel"], d["clocks"])
except Exception as e:
    print("bench ERR", e)
try:
    d = json.loads(open("gpurun_out/f_bench_magvit.json").read().strip().splitlines()[-1])
    print("magvit", round(d["encode_img_s"]), round(d["decode_img_s"]), round(d["encode_frac"], 3), round(d["decode_frac"], 3))
except Exception as e:
    print("magvit ERR", e)
PY
tail -c 400 gpurun_out/f_bench_reference.json
