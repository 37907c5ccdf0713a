#!/bin/bash
# MAGVIT2 follow-up (one gpurun call): tokenizer tests with the new defaults (32 images per pass, 8-row stem blocks),
# A/B of the stem variants and of 32 / 64 images per pass on one box, and the ncu launch list of one encode + decode pass.
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 600 python -m pytest -q -x -m gpu test_gpu_magvit.py test_gpu_cli.py > ../gpurun_out/c2_tests.log 2>&1; echo "magvit tests rc=$?" > ../gpurun_out/c2_summary.txt; cd ..
run() { # name, env
  env $2 timeout -k 10 200 python scripts/bench_magvit.py 64 >> gpurun_out/c2_magvit_$1.json 2>> gpurun_out/c2_magvit.err
  echo "magvit $1 rc=$?" >> gpurun_out/c2_summary.txt
}
for rep in 1 2; do
  run per32_stem1 "GENIE_B200_VQ_PER=32 GENIE_B200_STEM_ROWS=1"
  run per32_stem8 "GENIE_B200_VQ_PER=32"
  run per64_stem8 "GENIE_B200_VQ_PER=64"
done
timeout -k 10 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/c2_magvit_launches.csv python scripts/magvit_one_pass.py 32 > gpurun_out/c2_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/c2_summary.txt
python scripts/summarize_launches.py gpurun_out/c2_magvit_launches.csv "MAGVIT2 one encode + one decode pass, 32 images, fp16" > gpurun_out/c2_magvit_launch_shares.md 2>&1
cat gpurun_out/c2_summary.txt; tail -3 gpurun_out/c2_tests.log
for n in per32_stem1 per32_stem8 per64_stem8; do python - <<PY
import json
for l in open("gpurun_out/c2_magvit_$n.json"):
    d = json.loads(l); print("$n", round(d["encode_img_s"]), round(d["decode_img_s"]), round(d["encode_frac"], 3), round(d["decode_frac"], 3))
PY
done
head -30 gpurun_out/c2_magvit_launch_shares.md
