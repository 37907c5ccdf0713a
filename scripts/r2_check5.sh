#!/bin/bash
# round-2 GPU run 5: magvit tests after the ticket fix, lanes x SM-split experiment, magvit bench
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 900 python -m pytest -q -x -rP -m gpu test_gpu_magvit.py test_gpu_model.py -k "magvit or precision_modes or decode or encode or roundtrip or pipeline or lanes or fallback" > ../gpurun_out/r2_tests5.log 2>&1; echo "tests rc=$?"; cd ..
tail -3 gpurun_out/r2_tests5.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-parity --eval-clips 0"
timeout -k 10 300 $B > gpurun_out/r2_ab_lanes1.json 2> gpurun_out/r2_ab_lanes1.err; echo "lanes1 rc=$?"
GENIE_B200_SM_DIV=2 timeout -k 10 300 $B --lanes 2 > gpurun_out/r2_ab_lanes2_div2.json 2> gpurun_out/r2_ab_lanes2_div2.err; echo "lanes2 div2 rc=$?"
timeout -k 10 300 $B --lanes 2 > gpurun_out/r2_ab_lanes2_div1.json 2> gpurun_out/r2_ab_lanes2_div1.err; echo "lanes2 div1 rc=$?"
GENIE_B200_SM_DIV=2 timeout -k 10 300 $B --lanes 2 --chunk-tokens 16384 > gpurun_out/r2_ab_lanes2_div2_c16k.json 2> gpurun_out/r2_ab_lanes2_div2_c16k.err; echo "lanes2 div2 c16k rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_ab_lanes*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY
GENIE_PRECISION=fp16 timeout -k 10 300 python scripts/bench_magvit.py 64 > gpurun_out/r2_bench_magvit_fp16b.json 2> gpurun_out/r2_bench_magvit_fp16b.err; cat gpurun_out/r2_bench_magvit_fp16b.json
