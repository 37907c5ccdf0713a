#!/bin/bash
# MAGVIT2: CTA-pair convolution tiles (GENIE_B200_CONV_PAIR=1) and the mma.sync output conv (GENIE_B200_OUT_CONV_MMA=1):
# tokenizer tests with both switches on, then A/B on one box and the launch list with both on.
set -u
mkdir -p gpurun_out
cd tests
GENIE_B200_CONV_PAIR=1 GENIE_B200_OUT_CONV_MMA=1 timeout -k 10 300 python -m pytest -q -x -rP -m gpu test_gpu_magvit.py > ../gpurun_out/c3_tests_both.log 2>&1; echo "magvit tests (pair + mma) rc=$?" > ../gpurun_out/c3_summary.txt
GENIE_B200_OUT_CONV_MMA=1 timeout -k 10 300 python -m pytest -q -x -rP -m gpu test_gpu_magvit.py > ../gpurun_out/c3_tests_mma.log 2>&1; echo "magvit tests (mma only) rc=$?" >> ../gpurun_out/c3_summary.txt
cd ..
run() { # name, env
  env $2 timeout -k 10 200 python scripts/bench_magvit.py 64 >> gpurun_out/c3_magvit_$1.json 2>> gpurun_out/c3_magvit.err
  echo "magvit $1 rc=$?" >> gpurun_out/c3_summary.txt
}
for rep in 1 2; do
  run base "GENIE_B200_CONV_PAIR=0 GENIE_B200_OUT_CONV_MMA=0"
  run pair "GENIE_B200_CONV_PAIR=1 GENIE_B200_OUT_CONV_MMA=0"
  run mma "GENIE_B200_CONV_PAIR=0 GENIE_B200_OUT_CONV_MMA=1"
  run both "GENIE_B200_CONV_PAIR=1 GENIE_B200_OUT_CONV_MMA=1"
done
GENIE_B200_CONV_PAIR=1 GENIE_B200_OUT_CONV_MMA=1 timeout -k 10 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/c3_magvit_launches_both.csv python scripts/magvit_one_pass.py 32 > gpurun_out/c3_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/c3_summary.txt
python scripts/summarize_launches.py gpurun_out/c3_magvit_launches_both.csv "MAGVIT2 one encode + one decode pass, 32 images, fp16, CTA-pair conv tiles + mma.sync output conv" > gpurun_out/c3_magvit_launch_shares_both.md 2>&1
cat gpurun_out/c3_summary.txt; tail -4 gpurun_out/c3_tests_both.log; tail -4 gpurun_out/c3_tests_mma.log; grep -h "magvit fp16\|magvit bf16\|decode rel" gpurun_out/c3_tests_mma.log | head
for n in base pair mma both; do python - <<PY
import json
for l in open("gpurun_out/c3_magvit_$n.json"):
    d = json.loads(l); print("$n", round(d["encode_img_s"]), round(d["decode_img_s"]), round(d["encode_frac"], 3), round(d["decode_frac"], 3))
PY
done
head -24 gpurun_out/c3_magvit_launch_shares_both.md
