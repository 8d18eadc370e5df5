#!/bin/bash
# ncu evidence for the bench command of the final build: launch list (search + exact kernels; the graph build's ~3000 launches are
# filtered out) and one --set full capture of the traversal kernel on the bench's own graph (DRAM traffic per launch)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
K='regex:hnsw_search_kernel|hnsw_search_warp_kernel|exact_|merge_topk|to_half|col_bias|fill_empty|sql_|query_slack'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r02_launches_bench3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench3.log 2>&1; tail -1 gpurun_out/r02_ncu_bench3.log | cut -c1-160
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 8 -c 1 -o gpurun_out/r02_prof_search_bench \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_search_bench.log 2>&1; tail -2 gpurun_out/r02_ncu_search_bench.log | cut -c1-160
