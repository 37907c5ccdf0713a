#!/bin/bash
# A/B runs of bench.py on ONE box (SM clocks under the power cap differ box to box, so only same-box numbers compare):
#   scripts/ab_bench.sh name1 "ENV1=a ENV2=b" name2 "ENV1=c" ...      (use "X=1" for the unmodified baseline)
# prints frames/s, ms per step and the per-category kernel times of each variant.
set -u
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  name=$1; envs=$2; shift 2
  env $envs timeout -k 10 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"], 1), "frames/s", round(d["ms_per_step"], 1), "ms",
          {k: round(v["ms_per_step"], 1) for k, v in d["roofline"]["kernel_ms_by_category"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name", "ERR", e)
PY
done
