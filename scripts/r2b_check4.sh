#!/bin/bash
# MAGVIT2: persistent fused GroupNorm kernel (GENIE_B200_GN_FUSED=1): tokenizer tests with the switch on, A/B, launch list.
set -u
mkdir -p gpurun_out
cd tests
GENIE_B200_GN_FUSED=1 timeout -k 10 240 python -m pytest -q -x -rP -m gpu test_gpu_magvit.py > ../gpurun_out/c4_tests_fused.log 2>&1; echo "magvit tests (gn fused) rc=$?" > ../gpurun_out/c4_summary.txt
cd ..
run() { # name, env
  env $2 timeout -k 10 120 python scripts/bench_magvit.py 64 >> gpurun_out/c4_magvit_$1.json 2>> gpurun_out/c4_magvit.err
  echo "magvit $1 rc=$?" >> gpurun_out/c4_summary.txt
}
for rep in 1 2; do
  run base "GENIE_B200_GN_FUSED=0"
  run fused "GENIE_B200_GN_FUSED=1"
done
GENIE_B200_GN_FUSED=1 timeout -k 10 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/c4_magvit_launches_fused.csv python scripts/magvit_one_pass.py 32 > gpurun_out/c4_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/c4_summary.txt
python scripts/summarize_launches.py gpurun_out/c4_magvit_launches_fused.csv "MAGVIT2 one encode + one decode pass, 32 images, fp16, fused GroupNorm kernel + mma.sync output conv" > gpurun_out/c4_magvit_launch_shares_fused.md 2>&1
cat gpurun_out/c4_summary.txt; tail -4 gpurun_out/c4_tests_fused.log; grep -h "magvit fp16\|magvit bf16\|magvit fp32" gpurun_out/c4_tests_fused.log | head
for n in base fused; do python - <<PY
import json
for l in open("gpurun_out/c4_magvit_$n.json"):
    d = json.loads(l); print("$n", round(d["encode_img_s"]), round(d["decode_img_s"]), round(d["encode_frac"], 3), round(d["decode_frac"], 3))
PY
done
head -22 gpurun_out/c4_magvit_launch_shares_fused.md
