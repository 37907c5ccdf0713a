#!/bin/bash
# round-2 GPU run 3: full suite with per-test output, reduce-add epilogue A/B, tile-shape sweep for the N=512 GEMMs
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 1200 python -m pytest -q -x -rP -m gpu . > ../gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?"; cd ..
tail -3 gpurun_out/r2_tests3.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --eval-clips 0"
for red in 0 1; do
  GENIE_B200_RED_EPI=$red timeout -k 10 300 $B > gpurun_out/r2_ab_red$red.json 2> gpurun_out/r2_ab_red$red.err; echo "red=$red rc=$?"
done
for bn in 64 128 256 1128 1256; do
  GENIE_B200_RED_EPI=1 GENIE_B200_BN512=$bn timeout -k 10 300 $B > gpurun_out/r2_ab_red1_bn$bn.json 2> gpurun_out/r2_ab_red1_bn$bn.err; echo "red=1 bn=$bn rc=$?"
done
for bn in 64 128; do
  GENIE_B200_RED_EPI=0 GENIE_B200_BN512=$bn timeout -k 10 300 $B > gpurun_out/r2_ab_red0_bn$bn.json 2> gpurun_out/r2_ab_red0_bn$bn.err; echo "red=0 bn=$bn rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_ab_red*.json")):
    try:
        d = json.load(open(f))
        c = d["roofline"]["kernel_ms_by_category"]
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), {k: round(v["ms_per_step"], 1) for k, v in c.items()}, d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY
