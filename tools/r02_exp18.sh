#!/bin/bash
# exact path: warp-converged producer / MMA issuer (uniform-register descriptors) — parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py"
timeout 900 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest18.log | cut -c1-200
TURDB_EXACT_PAIR=1 timeout 600 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest18_pair.log 2>&1; echo "pytest pair rc=$?"; tail -3 gpurun_out/r02_pytest18_pair.log | cut -c1-200
for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 128 --metric 0 --gen sift_like" "--dim 128 --metric 2 --gen gaussian_latent" "--dim 768 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S"; timeout 200 python tools/exact_probe.py $SH --debug --out gpurun_out/r02_exact10.$S.json > gpurun_out/r02_exact10.$S.log 2>&1; tail -2 gpurun_out/r02_exact10.$S.log | cut -c1-20,180-420
done
echo "== pair 384 / 128ip"
TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py --dim 384 --metric 1 --gen gaussian_latent --out gpurun_out/r02_exact10_pair384.json 2>&1 | tail -1 | cut -c1-20,180-420
TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py --dim 128 --metric 2 --gen gaussian_latent --out gpurun_out/r02_exact10_pair128ip.json 2>&1 | tail -1 | cut -c1-20,180-420
