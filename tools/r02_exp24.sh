#!/bin/bash
# exact path: two alternating sets of epilogue warps (set s serves accumulator s), epilogue-bias form removed — parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py"
timeout 900 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest24.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest24.log | cut -c1-200
TURDB_EXACT_PAIR=0 timeout 600 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest24_single.log 2>&1; echo "pytest one-CTA rc=$?"; tail -2 gpurun_out/r02_pytest24_single.log | cut -c1-200
for SH in "--dim 128 --metric 0 --gen sift_like" "--dim 128 --metric 2 --gen gaussian_latent" "--dim 64 --metric 2 --gen gaussian_latent" "--dim 384 --metric 1 --gen gaussian_latent" "--dim 768 --metric 2 --gen gaussian_latent" "--dim 256 --metric 1 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S"; TURDB_EXACT_VERBOSE=1 timeout 200 python tools/exact_probe.py $SH --debug --out gpurun_out/r02_exact15.$S.json > gpurun_out/r02_exact15.$S.log 2>&1; grep -m1 "turdb exact" gpurun_out/r02_exact15.$S.log; grep -h "dbg cycles" gpurun_out/r02_exact15.$S.log | cut -c1-250; tail -1 gpurun_out/r02_exact15.$S.log | cut -c1-20,180-420
done
