#!/bin/bash
# exact path: warp-per-query radix-select threshold kernel (parity + timing), slice growth 2/3/4, ncu --set full of the last pass
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="tests/test_gpu_exact.py tests/test_gpu_sql_operator.py"
timeout 600 python -m pytest $T -m gpu -q -x > gpurun_out/r02_pytest13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest13.log | cut -c1-200
SH="--dim 384 --metric 1 --gen gaussian_latent"
for G in 2 3 4; do
  echo "== growth $G"; TURDB_EXACT_GROWTH=$G timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact5_growth$G.json 2>&1 | tail -1 | cut -c180-330
done
echo "== pair"; TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact5_pair.json 2>&1 | tail -1 | cut -c180-330
echo "== 128-d, 768-d"
timeout 200 python tools/exact_probe.py --dim 128 --metric 0 --gen sift_like --out gpurun_out/r02_exact5_128.json 2>&1 | tail -1 | cut -c180-330
timeout 200 python tools/exact_probe.py --dim 768 --metric 2 --gen gaussian_latent --out gpurun_out/r02_exact5_768.json 2>&1 | tail -1 | cut -c180-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'exact_|to_half|query_slack|col_bias|fill_' -c 200 --csv --log-file gpurun_out/r02_launches_exact_probe2.csv \
   python tools/exact_probe.py $SH --reps 1 --out gpurun_out/ncu_dummy.json > gpurun_out/r02_ncu_exact_list2.log 2>&1; tail -1 gpurun_out/r02_ncu_exact_list2.log | cut -c1-100
timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_gemm_filter_kernel -s 17 -c 1 -o gpurun_out/r02_prof_exact_epi16 \
    python tools/exact_probe.py $SH --reps 2 --out gpurun_out/ncu_dummy.json > gpurun_out/r02_ncu_exact2.log 2>&1; tail -2 gpurun_out/r02_ncu_exact2.log | cut -c1-200
