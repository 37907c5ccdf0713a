set -u
for d in 0 16; do
  echo "DBG=$d"
  for sh in "proj+res(dual)" "proj+res" "fc2+res"; do
  GENIE_B200_GEMM_DEBUG=$d timeout -k 5 200 python scripts/gemm_microbench.py "$sh" 32768
  done
  GENIE_B200_GEMM_DEBUG=$d timeout -k 5 200 python scripts/gemm_microbench.py "proj+res(dual)" 262144
done
cd tests; timeout 300 python -m pytest -q -x -m gpu test_gpu_kernels.py -k linear 2>&1 | tail -3
