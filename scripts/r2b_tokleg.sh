#!/bin/bash
# quick check of bench.py's tokenizer leg in isolation (non-default stream, like bench.py's main)
mkdir -p gpurun_out
timeout -k 5 90 python - > gpurun_out/t_tokleg.json 2> gpurun_out/t_tokleg.err <<'PY'
import importlib, json, sys, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream(dev))
pkg = importlib.import_module("1xgpt_b200")
print(json.dumps(bench.tokenizer_leg(pkg, dev, 1401.5)))
PY
echo "rc=$?"; cat gpurun_out/t_tokleg.json; tail -3 gpurun_out/t_tokleg.err
