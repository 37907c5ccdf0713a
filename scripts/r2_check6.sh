#!/bin/bash
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 600 python -m pytest -q -x -rP -m gpu test_gpu_model.py -k "fused_readout or maskgit_argmax or teacher_forced_eval_138m" > ../gpurun_out/r2_tests6.log 2>&1; echo "tests rc=$?"; cd ..
tail -3 gpurun_out/r2_tests6.log; grep -h "fused vs" gpurun_out/r2_tests6.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --eval-clips 0"
for f in 1 0; do GENIE_B200_FUSED_READOUT=$f timeout -k 10 300 $B > gpurun_out/r2_ab_fro$f.json 2> gpurun_out/r2_ab_fro$f.err; echo "fro=$f rc=$?"; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_ab_fro*.json")):
    d = json.load(open(f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 2), d["gpu_launches"], d["clocks"]["sm_mhz"])
PY
