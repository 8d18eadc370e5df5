#!/bin/bash
# exact path: 256-column tiles + 8 epilogue warps vs the 128-column build; then the N = 1 points of configs 5 and 4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sql_operator.py tests/test_gpu_build.py -m gpu -q -x > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest8.log | cut -c1-200
for LIB in libturdb_cuda.so libturdb_cuda.t128.so; do
  for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 128 --metric 0 --gen sift_like" "--dim 768 --metric 2 --gen gaussian_latent"; do
    TURDB_CUDA_LIB=$PWD/turdb_b200/$LIB timeout 300 python tools/exact_probe.py $SH --out gpurun_out/r02_exact_$LIB.$(echo $SH | tr -d ' -').json 2>&1 | tail -2 | cut -c1-300
  done
done
timeout 1200 python tools/run_sharded.py --config 5 --rows-per-rank 12500000 --out gpurun_out/r02_config5_n1.json > gpurun_out/r02_config5_n1.log 2>&1; echo "c5 rc=$?"
grep -h "^{\"ef" gpurun_out/r02_config5_n1.log | cut -c1-300
timeout 2400 python tools/run_sharded.py --config 4 --total-rows 10000000 --out gpurun_out/r02_config4_n1.json > gpurun_out/r02_config4_n1.log 2>&1; echo "c4 rc=$?"
grep -h "^{\"ef" gpurun_out/r02_config4_n1.log | cut -c1-300; tail -2 gpurun_out/r02_config4_n1.log | cut -c1-300
