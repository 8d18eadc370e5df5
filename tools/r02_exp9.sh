#!/bin/bash
# 8 GPUs: config 5 again with the small ef values (the first ef with recall >= 0.95 is below 64 on this graph)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 1200 $TR tools/run_sharded.py --config 5 --rows-per-rank 12500000 --efs 8,16,32,64 --parity 0 --out gpurun_out/r02_config5_n8_small_ef.json > gpurun_out/r02_config5_n8_small_ef.log 2>&1; echo "c5 rc=$?"
grep -h "^{\"ef" gpurun_out/r02_config5_n8_small_ef.log | cut -c1-330
