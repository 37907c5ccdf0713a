#!/bin/bash
# Second-session validation (one gpurun call): full GPU suite (incl. the new evaluator / eval_utils tests), smoke, and the
# MAGVIT2 images-per-pass A/B (GENIE_B200_VQ_PER = 8 / 16 / 32 on one box).
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 900 python -m pytest -q -x -m gpu . > ../gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?" > ../gpurun_out/c1_summary.txt; cd ..
timeout -k 10 200 python __graft_entry__.py smoke > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c1_summary.txt
for per in 8 16 32 8 16 32; do
  GENIE_B200_VQ_PER=$per timeout -k 10 200 python scripts/bench_magvit.py 64 >> gpurun_out/c1_magvit_per$per.json 2>> gpurun_out/c1_magvit.err
  echo "magvit per=$per rc=$?" >> gpurun_out/c1_summary.txt
done
cat gpurun_out/c1_summary.txt; tail -3 gpurun_out/c1_tests.log
for per in 8 16 32; do python - <<PY
import json
for l in open("gpurun_out/c1_magvit_per$per.json"):
    d = json.loads(l); print("per=$per", round(d["encode_img_s"]), round(d["decode_img_s"]), round(d["encode_frac"], 3), round(d["decode_frac"], 3))
PY
done
