#!/bin/bash
# round-2 final validation on one GPU: all GPU tests, smoke, both bench arms, ncu launch list of the bench, ncu --set full of the
# exact path's two-CTA kernel (largest pass at 1M x 384)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_final2.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference2.log 2>&1; tail -1 gpurun_out/r02_bench_reference2.log | cut -c1-300
timeout 900 python bench.py --out gpurun_out/r02_bench_final2_n1.json > gpurun_out/r02_bench_final2_n1.log 2>&1; tail -1 gpurun_out/r02_bench_final2_n1.log | cut -c1-400
K='regex:hnsw_search|exact_|merge_topk|sanitize_adj|norm2_kernel|to_half|col_bias|fill_empty|sql_|insert_|bf16_rowerr|query_slack|max_abs|DeviceRadixSort'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 3000 --csv --log-file gpurun_out/r02_launches_bench2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench2.log 2>&1; tail -1 gpurun_out/r02_ncu_bench2.log | cut -c1-200
SH="--dim 384 --metric 1 --gen gaussian_latent"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_gemm_filter_pair_kernel -s 22 -c 1 -o gpurun_out/r02_prof_exact_pair \
    python tools/exact_probe.py $SH --reps 2 --out gpurun_out/ncu_dummy.json > gpurun_out/r02_ncu_exact_pair.log 2>&1; tail -2 gpurun_out/r02_ncu_exact_pair.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'exact_|to_half|query_slack|col_bias|fill_' -c 200 --csv --log-file gpurun_out/r02_launches_exact_probe3.csv \
   python tools/exact_probe.py $SH --reps 1 --out gpurun_out/ncu_dummy.json > /dev/null 2>&1; echo list done
