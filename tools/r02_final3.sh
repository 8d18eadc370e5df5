#!/bin/bash
# round-2 final validation (3): all GPU tests in the default configuration, exact/SQL tests again in the one-CTA form, smoke, both
# bench arms, L2 / IP probes of the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_final3.log | cut -c1-200
TURDB_EXACT_PAIR=0 timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sql_operator.py -m gpu -q -x > gpurun_out/r02_pytest_final3_single.log 2>&1; echo "pytest one-CTA rc=$?"; tail -2 gpurun_out/r02_pytest_final3_single.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference3.log 2>&1; tail -1 gpurun_out/r02_bench_reference3.log | cut -c1-200
timeout 900 python bench.py --out gpurun_out/r02_bench_final3_n1.json > gpurun_out/r02_bench_final3_n1.log 2>&1; tail -1 gpurun_out/r02_bench_final3_n1.log | cut -c1-300
for SH in "--dim 128 --metric 0 --gen sift_like" "--dim 384 --metric 0 --gen gaussian_latent" "--dim 384 --metric 1 --gen gaussian_latent" "--dim 512 --metric 2 --gen gaussian_latent" "--dim 64 --metric 2 --gen gaussian_latent"; do
  S=$(echo $SH | tr -d ' -')
  echo "== $S"; TURDB_EXACT_VERBOSE=1 timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact14.$S.json > gpurun_out/r02_exact14.$S.log 2>&1; grep -m1 "turdb exact" gpurun_out/r02_exact14.$S.log; tail -1 gpurun_out/r02_exact14.$S.log | cut -c1-20,180-420
done
