#!/bin/bash
# exact path with FP16 operands (vs forced BF16, vs no slack band), form calibration at ~14 rows/hop, ncu captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sql_operator.py tests/test_gpu_configs_scaled.py -m gpu -q -x > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest10.log | cut -c1-200
for SH in "--dim 384 --metric 1 --gen gaussian_latent" "--dim 128 --metric 0 --gen sift_like" "--dim 768 --metric 2 --gen gaussian_latent" "--dim 128 --metric 0 --gen gaussian_latent"; do
  T=$(echo $SH | tr -d ' -')
  timeout 300 python tools/exact_probe.py $SH --out gpurun_out/r02_exact2_fp16.$T.json 2>&1 | tail -1 | cut -c1-60,330-520
  TURDB_EXACT_FORCE_BF16=1 timeout 300 python tools/exact_probe.py $SH --out gpurun_out/r02_exact2_bf16.$T.json 2>&1 | tail -1 | cut -c1-60,330-520
  TURDB_EXACT_SLACK_SCALE=0 timeout 300 python tools/exact_probe.py $SH --out gpurun_out/r02_exact2_noslack.$T.json 2>&1 | tail -1 | cut -c1-60,330-520
done
timeout 900 python tools/sweep.py --n 4000000 --dim 128 --metric 0 --gen clustered --genkw '{"centre_latent":16,"sigma":0.3,"corpus_n":32000000}' --builder insert --ef 64 --tunings "0,0,0,0,1;0,0,0,0,2" --out gpurun_out/r02_e10_form_calib.json > gpurun_out/r02_e10_form_calib.log 2>&1
grep -h "^{" gpurun_out/r02_e10_form_calib.log | cut -c1-220
# ncu: the exact kernel (largest pass), then the direct traversal kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_gemm_filter_kernel -s 17 -c 1 -o gpurun_out/r02_prof_exact_tile256 \
    python tools/exact_probe.py --dim 384 --metric 1 --gen gaussian_latent --reps 2 --out gpurun_out/ncu_exact_dummy.json > gpurun_out/r02_ncu_exact.log 2>&1; tail -2 gpurun_out/r02_ncu_exact.log | cut -c1-200
A="2000000 --dim 128 --metric 0 --gen clustered --genkw {\"centre_latent\":16,\"sigma\":0.3} --ef 128"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_warp_kernel -s 6 -c 1 -o gpurun_out/r02_prof_direct_final \
    python tools/ncu_target.py $A --reps 3 > gpurun_out/r02_ncu_direct.log 2>&1; tail -2 gpurun_out/r02_ncu_direct.log | cut -c1-200
