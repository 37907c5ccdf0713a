#!/bin/bash
# CTA-pair convolution tiles (GENIE_B200_CONV_PAIR=1): tokenizer tests with the switch on, then A/B on one box.
set -u
mkdir -p gpurun_out
cd tests && GENIE_B200_CONV_PAIR=1 timeout -k 10 300 python -m pytest -q -x -m gpu test_gpu_magvit.py > ../gpurun_out/c3_tests.log 2>&1; echo "magvit tests (conv pair) rc=$?" > ../gpurun_out/c3_summary.txt; cd ..
run() { # name, env
  env $2 timeout -k 10 200 python scripts/bench_magvit.py 64 >> gpurun_out/c3_magvit_$1.json 2>> gpurun_out/c3_magvit.err
  echo "magvit $1 rc=$?" >> gpurun_out/c3_summary.txt
}
for rep in 1 2; do
  run single "GENIE_B200_CONV_PAIR=0"
  run pair "GENIE_B200_CONV_PAIR=1"
done
GENIE_B200_CONV_PAIR=1 timeout -k 10 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/c3_magvit_launches_pair.csv python scripts/magvit_one_pass.py 32 > gpurun_out/c3_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/c3_summary.txt
python scripts/summarize_launches.py gpurun_out/c3_magvit_launches_pair.csv "MAGVIT2 one encode + one decode pass, 32 images, fp16, CTA-pair conv tiles" > gpurun_out/c3_magvit_launch_shares_pair.md 2>&1
cat gpurun_out/c3_summary.txt; tail -5 gpurun_out/c3_tests.log
for n in single pair; do python - <<PY
import json
for l in open("gpurun_out/c3_magvit_$n.json"):
    d = json.loads(l); print("$n", round(d["encode_img_s"]), round(d["decode_img_s"]), round(d["encode_frac"], 3), round(d["decode_frac"], 3))
PY
done
head -24 gpurun_out/c3_magvit_launch_shares_pair.md
