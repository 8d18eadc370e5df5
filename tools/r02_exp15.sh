#!/bin/bash
# exact path: dense first slice, bias prefetch, longer items at short K, k' = k default, growth 3 — parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py tests/test_gpu_configs_scaled.py tests/test_hnsw_file.py"
timeout 900 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest15.log | cut -c1-200
TURDB_EXACT_PAIR=1 timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sql_operator.py -m gpu -q -x > gpurun_out/r02_pytest15_pair.log 2>&1; echo "pytest pair rc=$?"; tail -3 gpurun_out/r02_pytest15_pair.log | cut -c1-200
for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 128 --metric 0 --gen sift_like" "--dim 768 --metric 2 --gen gaussian_latent" "--dim 128 --metric 0 --gen gaussian_latent" "--dim 128 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S"; timeout 200 python tools/exact_probe.py $SH --debug --out gpurun_out/r02_exact7.$S.json 2>&1 | tail -2 | cut -c1-60,180-360
done
