#!/bin/bash
# round-2 GPU run 4: full suite (hd32 tcgen05 spatial, MAGVIT2 precision modes, CLI, pipeline), 35M / 700M bench lines
set -u
mkdir -p gpurun_out
cd tests && timeout -k 10 1500 python -m pytest -q -x -rP -m gpu . > ../gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?"; cd ..
tail -4 gpurun_out/r2_tests4.log
grep -h "magvit fp\|magvit bf\|spatial attention hd\|fp16 linear" gpurun_out/r2_tests4.log
timeout -k 10 300 python scripts/bench_generate.py --layers 32 --d-model 256 --heads 8 --batch 64 --maskgit-steps 2 > gpurun_out/r2_bench_35m.json 2> gpurun_out/r2_bench_35m.err; echo "35m rc=$?"; cat gpurun_out/r2_bench_35m.json
GENIE_B200_SPATIAL_TC=0 timeout -k 10 300 python scripts/bench_generate.py --layers 32 --d-model 256 --heads 8 --batch 64 --maskgit-steps 2 > gpurun_out/r2_bench_35m_mma.json 2>> gpurun_out/r2_bench_35m.err; echo "35m mma rc=$?"; cat gpurun_out/r2_bench_35m_mma.json
for p in fp16 bf16; do GENIE_PRECISION=$p timeout -k 10 300 python scripts/bench_magvit.py 64 > gpurun_out/r2_bench_magvit_$p.json 2> gpurun_out/r2_bench_magvit_$p.err; echo "magvit $p rc=$?"; cat gpurun_out/r2_bench_magvit_$p.json; done
