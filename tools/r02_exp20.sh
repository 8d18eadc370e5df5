#!/bin/bash
# exact path, L2: the bias inside the contraction (three augmented columns) against the epilogue bias add — parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py tests/test_gpu_configs_scaled.py tests/test_hnsw_file.py"
timeout 900 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest20.log | cut -c1-200
TURDB_EXACT_FORCE_BF16=1 timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sql_operator.py -m gpu -q -x > gpurun_out/r02_pytest20_bf16.log 2>&1; echo "pytest bf16 rc=$?"; tail -3 gpurun_out/r02_pytest20_bf16.log | cut -c1-200
for SH in "--dim 128 --metric 0 --gen sift_like" "--dim 128 --metric 0 --gen gaussian_latent" "--dim 384 --metric 0 --gen gaussian_latent" "--dim 768 --metric 0 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  for A in 1 0; do
    echo "== $S aug=$A"; TURDB_EXACT_L2_AUG=$A TURDB_EXACT_VERBOSE=1 timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact12_aug$A.$S.json > gpurun_out/r02_exact12_aug$A.$S.log 2>&1; grep -m1 "turdb exact" gpurun_out/r02_exact12_aug$A.$S.log; tail -1 gpurun_out/r02_exact12_aug$A.$S.log | cut -c1-20,180-420
  done
done
