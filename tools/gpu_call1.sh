#!/bin/bash
# round-1 measurement call: tests, bench, ncu launch list of the bench command, ncu --set full of the traversal kernel, tuning sweeps
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --out gpurun_out/bench_r01_n1.json > gpurun_out/bench_default.log 2>&1
K='regex:hnsw_search_kernel|exact_|merge_topk|sanitize_adj|norm2_kernel|to_bf16|col_bias|fill_empty'
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/ncu_target.py > gpurun_out/plain_target.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -o gpurun_out/prof_search_team \
    python tools/ncu_target.py > gpurun_out/ncu_target.log 2>&1
timeout 600 python tools/sweep.py --debug --tunings "0,0,0,0;4,16,0,1;4,24,0,1;4,32,0,1;2,16,0,1;3,16,0,1" \
    --probe "5,16,46080;5,16,0;8,16,0;4,32,0;8,32,0;2,64,0;1,128,0" --out gpurun_out/sweep_384.json > gpurun_out/sweep_384.log 2>&1
timeout 600 python tools/sweep.py --debug --dim 128 --metric 0 --gen sift_like --tunings "0,0,0,0;4,16,0,1;4,32,0,1;2,16,0,1;2,8,0,1;2,32,0,1;1,8,0,1;1,16,0,1" \
    --probe "5,16,0;8,16,0;8,32,0;4,64,0;16,16,0" --out gpurun_out/sweep_128.json > gpurun_out/sweep_128.log 2>&1
echo done
