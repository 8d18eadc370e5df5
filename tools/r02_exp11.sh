#!/bin/bash
# exact path: 16-warp vote/queue epilogue (1-CTA) and the two-CTA (cta_group::2) form; parity tests in both, probes with per-role cycle counters
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py"
timeout 600 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest11_single.log 2>&1; echo "pytest single rc=$?"; tail -3 gpurun_out/r02_pytest11_single.log | cut -c1-200
TURDB_EXACT_PAIR=1 timeout 400 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest11_pair.log 2>&1; echo "pytest pair rc=$?"; tail -3 gpurun_out/r02_pytest11_pair.log | cut -c1-200
for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 128 --metric 0 --gen sift_like" "--dim 768 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  for P in 0 1; do
    echo "== $S pair=$P"
    TURDB_EXACT_PAIR=$P timeout 200 python tools/exact_probe.py $SH --debug --out gpurun_out/r02_exact3_pair$P.$S.json 2>&1 | tail -2 | cut -c1-330
  done
done
for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 128 --metric 0 --gen sift_like"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S epi128 (8 epilogue warps)"
  TURDB_CUDA_LIB=$PWD/turdb_b200/libturdb_cuda.epi128.so timeout 200 python tools/exact_probe.py $SH --debug --out gpurun_out/r02_exact3_epi128.$S.json 2>&1 | tail -2 | cut -c1-330
done
