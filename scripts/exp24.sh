set -u
mkdir -p gpurun_out
cd tests; timeout 600 python -m pytest -q -x -m gpu test_gpu_magvit.py -s 2>&1 | tail -15; cd ..
timeout 200 python scripts/bench_magvit.py 64 2>&1 | tail -2
