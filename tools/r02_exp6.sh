#!/bin/bash
# N GPUs (default 8): config 5 (weak, 12.5M x 128 per rank) and config 4 (strong, 10M x 768 split over the ranks)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 1200 $TR tools/run_sharded.py --config 5 --rows-per-rank 12500000 --out gpurun_out/r02_config5_n$N.json > gpurun_out/r02_config5_n$N.log 2>&1; echo "c5 rc=$?"
grep -h "^{" gpurun_out/r02_config5_n$N.log | cut -c1-700 | tail -5
timeout 1200 $TR tools/run_sharded.py --config 4 --total-rows 10000000 --out gpurun_out/r02_config4_n$N.json > gpurun_out/r02_config4_n$N.log 2>&1; echo "c4 rc=$?"
grep -h "^{" gpurun_out/r02_config4_n$N.log | cut -c1-700 | tail -3
nvidia-smi --query-gpu=index,name,memory.used --format=csv | head -10
