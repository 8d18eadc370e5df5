#!/bin/bash
# round-2 final validation (4): the shipped build — all GPU tests, smoke, bench (no CPU baseline), three probes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_final4.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --out gpurun_out/r02_bench_final4_n1.json > gpurun_out/r02_bench_final4_n1.log 2>&1; tail -1 gpurun_out/r02_bench_final4_n1.log | cut -c1-200
for SH in "--dim 128 --metric 0 --gen sift_like" "--dim 384 --metric 1 --gen gaussian_latent" "--dim 768 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S"; timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact16.$S.json 2>&1 | tail -1 | cut -c1-20,180-420
done
