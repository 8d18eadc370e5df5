#!/bin/bash
# exact path: 128-vector tiles over four accumulators + two alternating epilogue sets (short K) — parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact.py -m gpu -q -x > gpurun_out/r02_pytest27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest27.log | cut -c1-200
TURDB_EXACT_TILE_N=128 timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sql_operator.py -m gpu -q -x > gpurun_out/r02_pytest27_t128.log 2>&1; echo "pytest tile128 rc=$?"; tail -2 gpurun_out/r02_pytest27_t128.log | cut -c1-200
for SH in "--dim 128 --metric 0 --gen sift_like" "--dim 128 --metric 2 --gen gaussian_latent" "--dim 64 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S tile 128"; TURDB_EXACT_TILE_N=128 TURDB_EXACT_VERBOSE=1 timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact17_t128.$S.json > gpurun_out/r02_exact17_t128.$S.log 2>&1; grep -m1 "turdb exact" gpurun_out/r02_exact17_t128.$S.log; tail -1 gpurun_out/r02_exact17_t128.$S.log | cut -c1-20,180-420
done
echo "== 64 ip tile 128 pair"; TURDB_EXACT_TILE_N=128 TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py --dim 64 --metric 2 --gen gaussian_latent --out gpurun_out/r02_exact17_t128_pair64.json 2>&1 | tail -1 | cut -c1-20,180-420
