#!/bin/bash
# exact path: default form selection (two-CTA at 256..512 dims), parity in the default configuration, all shapes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py tests/test_gpu_configs_scaled.py"
timeout 900 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest19.log | cut -c1-200
for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 256 --metric 1 --gen gaussian_latent" "--dim 512 --metric 2 --gen gaussian_latent" "--dim 128 --metric 0 --gen sift_like" "--dim 768 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S"; TURDB_EXACT_VERBOSE=1 timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact11.$S.json > gpurun_out/r02_exact11.$S.log 2>&1; grep -m1 "turdb exact" gpurun_out/r02_exact11.$S.log; tail -1 gpurun_out/r02_exact11.$S.log | cut -c1-20,180-420
done
echo "== 256 single, 512 single, 768 pair"
TURDB_EXACT_PAIR=0 timeout 200 python tools/exact_probe.py --dim 256 --metric 1 --gen gaussian_latent --out gpurun_out/r02_exact11_single256.json 2>&1 | tail -1 | cut -c1-20,180-420
TURDB_EXACT_PAIR=0 timeout 200 python tools/exact_probe.py --dim 512 --metric 2 --gen gaussian_latent --out gpurun_out/r02_exact11_single512.json 2>&1 | tail -1 | cut -c1-20,180-420
TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py --dim 768 --metric 2 --gen gaussian_latent --out gpurun_out/r02_exact11_pair768.json 2>&1 | tail -1 | cut -c1-20,180-420
