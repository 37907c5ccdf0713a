set -u
mkdir -p gpurun_out
true
timeout 120 python scripts/spatial_microbench.py 64 8 20 >> gpurun_out/e11_micro.jsonl 2>&1
cat gpurun_out/e11_micro.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spatial_attn_tc -c 2 -o gpurun_out/e11_spatial_tc python scripts/spatial_microbench.py 128 8 1 0 > gpurun_out/e11_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/e11_ncu.log
