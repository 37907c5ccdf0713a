set -u
for pp in 1 0; do
  echo "PAIR_PROJ=$pp"
  GENIE_B200_PAIR_PROJ=$pp timeout -k 5 200 python scripts/gemm_microbench.py "proj+res(dual)" 32768
  GENIE_B200_PAIR_PROJ=$pp timeout -k 5 200 python scripts/gemm_microbench.py "proj+res" 32768
  GENIE_B200_PAIR_PROJ=$pp timeout -k 5 200 python scripts/gemm_microbench.py "proj+res" 16384
  GENIE_B200_PAIR_PROJ=$pp timeout -k 5 200 python scripts/gemm_microbench.py "proj+res(dual)" 262144
done
cd tests; timeout 300 python -m pytest -q -x -m gpu test_gpu_kernels.py -k linear 2>&1 | tail -3
