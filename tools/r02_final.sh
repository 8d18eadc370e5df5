#!/bin/bash
# round-2 final validation on one GPU: GPU tests, smoke, both bench arms, latent-32 variant, ncu launch list of the bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_final.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.log 2>&1; tail -1 gpurun_out/r02_bench_reference.log | cut -c1-700
timeout 900 python bench.py --out gpurun_out/r02_bench_final_n1.json > gpurun_out/r02_bench_final_n1.log 2>&1; tail -1 gpurun_out/r02_bench_final_n1.log | cut -c1-400
timeout 900 python bench.py --latent 32 --no-cpu-baseline --out gpurun_out/r02_bench_latent32.json > gpurun_out/r02_bench_latent32.log 2>&1; tail -1 gpurun_out/r02_bench_latent32.log | cut -c1-300
K='regex:hnsw_search|exact_|merge_topk|sanitize_adj|norm2_kernel|to_half|col_bias|fill_empty|sql_|insert_|bf16_rowerr|query_slack|max_abs|DeviceRadixSort'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; tail -1 gpurun_out/r02_ncu_bench.log | cut -c1-200
