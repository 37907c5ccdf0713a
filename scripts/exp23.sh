set -u
mkdir -p gpurun_out
timeout 200 python scripts/bench_magvit.py 64 > gpurun_out/e23_magvit.json 2> gpurun_out/e23_magvit.err; echo rc=$?; cat gpurun_out/e23_magvit.json
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/e23_magvit_launches.csv python scripts/bench_magvit.py 8 > gpurun_out/e23_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/e23_magvit_launches.csv
