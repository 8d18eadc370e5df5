#!/bin/bash
# N = $1 GPUs: config 5 (weak) and config 4 (strong) — the N < 8 points of the scaling curves
cd "$GRAFT_REPO_ROOT" || exit 1
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"; fi
timeout 1200 $TR tools/run_sharded.py --config 5 --rows-per-rank 12500000 --out gpurun_out/r02_config5_n$N.json > gpurun_out/r02_config5_n$N.log 2>&1; echo "c5 rc=$?"
grep -h "^{\"ef" gpurun_out/r02_config5_n$N.log | cut -c1-330
timeout 1800 $TR tools/run_sharded.py --config 4 --total-rows 10000000 --out gpurun_out/r02_config4_n$N.json > gpurun_out/r02_config4_n$N.log 2>&1; echo "c4 rc=$?"
grep -h "^{\"ef" gpurun_out/r02_config4_n$N.log | cut -c1-330; tail -3 gpurun_out/r02_config4_n$N.log | cut -c1-300
