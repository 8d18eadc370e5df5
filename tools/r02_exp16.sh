#!/bin/bash
# exact path: per-warp bias staging (no CTA barrier) — parity + timing; two-CTA form at 128-d (L2 -> SM bound there)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py"
timeout 900 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest16.log | cut -c1-200
for SH in "--dim 128 --metric 0 --gen sift_like" "--dim 128 --metric 2 --gen gaussian_latent" "--dim 384 --metric 0 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  for P in 0 1; do
    echo "== $S pair $P"; TURDB_EXACT_PAIR=$P timeout 200 python tools/exact_probe.py $SH --debug --out gpurun_out/r02_exact8_pair$P.$S.json 2>&1 | tail -2 | cut -c1-20,180-420
  done
done
