#!/bin/bash
# The measurement recipe behind profiles/r01_*: bench line, ncu launch list of the bench command, ncu --set full of the
# traversal kernel, single-GPU config runs.  Run on the GPU box: gpurun -- bash tools/profile_round.sh
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --out gpurun_out/bench_r01_n1.json > gpurun_out/bench_default.log 2>&1
K='regex:hnsw_search_kernel|exact_|merge_topk|sanitize_adj|norm2_kernel|to_bf16|col_bias|fill_empty|sql_rekey'
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/ncu_target.py > gpurun_out/plain_target.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -o gpurun_out/prof_search_r01 \
    python tools/ncu_target.py > gpurun_out/ncu_target.log 2>&1
timeout 900 python tools/run_configs.py --out gpurun_out/configs_r01.json > gpurun_out/configs.log 2>&1
echo done
