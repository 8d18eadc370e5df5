#!/bin/bash
# exact path, two-CTA form: CTA-scope remote arrive instead of .release.cluster — parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py"
TURDB_EXACT_PAIR=1 timeout 600 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest21_pair.log 2>&1; echo "pytest pair rc=$?"; tail -3 gpurun_out/r02_pytest21_pair.log | cut -c1-200
for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 256 --metric 1 --gen gaussian_latent" "--dim 768 --metric 2 --gen gaussian_latent" "--dim 128 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S pair"; TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact13_pair.$S.json 2>&1 | tail -1 | cut -c1-20,180-420
done
