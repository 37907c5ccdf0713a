#!/bin/bash
# A/B runs of the MAGVIT2 tokenizer on ONE box (SM clocks under the power cap differ box to box):
#   scripts/magvit_ab.sh name1 "ENV1=a ENV2=b" name2 "ENV1=c" ...       (use "X=1" for the unmodified defaults)
# per variant: the tokenizer parity tests with the switches on, two runs of scripts/bench_magvit.py (64 frames), and the ncu
# launch list of one encode + decode pass of 32 images (-> gpurun_out/ab_<name>_launch_shares.md).
# This is how profiles/r02b_* were produced (GENIE_B200_VQ_PER, _STEM_ROWS, _CONV_PAIR, _OUT_CONV_MMA, _GN_FUSED).
set -u
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  name=$1; envs=$2; shift 2
  (cd tests && env $envs timeout -k 10 300 python -m pytest -q -x -rP -m gpu test_gpu_magvit.py > ../gpurun_out/ab_${name}_tests.log 2>&1; echo "$name tests rc=$?")
  for rep in 1 2; do
    env $envs timeout -k 10 200 python scripts/bench_magvit.py 64 >> gpurun_out/ab_${name}.jsonl 2>> gpurun_out/ab_${name}.err
  done
  env $envs timeout -k 10 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/ab_${name}_launches.csv python scripts/magvit_one_pass.py 32 > gpurun_out/ab_${name}_ncu.log 2>&1
  python scripts/summarize_launches.py gpurun_out/ab_${name}_launches.csv "MAGVIT2 one encode + one decode pass, 32 images ($envs)" \
    > gpurun_out/ab_${name}_launch_shares.md 2>&1
  python - <<PY
import json
for l in open("gpurun_out/ab_${name}.jsonl"):
    d = json.loads(l)
    print("$name", round(d["encode_img_s"]), "/", round(d["decode_img_s"]), "img/s encode / decode,", round(d["encode_frac"], 3), "/", round(d["decode_frac"], 3), "of peak")
PY
done
