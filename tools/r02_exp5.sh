#!/bin/bash
# 2 GPUs: the sharded bench path (packed all-gather, pipelined e2e) and the config tool at a small size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --out gpurun_out/r02_bench_n2.json > gpurun_out/r02_bench_n2.log 2>&1; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_n2.log
timeout 900 $TR tools/run_sharded.py --config 5 --rows-per-rank 2000000 --out gpurun_out/r02_c5_n2_small.json > gpurun_out/r02_c5_n2_small.log 2>&1; echo "c5 rc=$?"
tail -c 2500 gpurun_out/r02_c5_n2_small.log
timeout 900 $TR tools/run_sharded.py --config 4 --total-rows 1000000 --out gpurun_out/r02_c4_n2_small.json > gpurun_out/r02_c4_n2_small.log 2>&1; echo "c4 rc=$?"
tail -c 2000 gpurun_out/r02_c4_n2_small.log
