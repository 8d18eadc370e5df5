#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
LIBS=turdb_b200/libturdb_cuda.tma.so,turdb_b200/libturdb_cuda.so,turdb_b200/libturdb_cuda.ldg_m1.so
timeout 900 python tools/sweep.py --debug --libs $LIBS --tunings "2,16,0,1;4,16,0,1;3,24,0,1;2,24,0,1;4,32,0,1" \
    --out gpurun_out/ab_384.json > gpurun_out/ab_384.log 2>&1
timeout 900 python tools/sweep.py --debug --dim 128 --metric 0 --gen sift_like --libs $LIBS \
    --tunings "2,16,0,1;1,16,0,1;4,16,0,1;2,24,0,1;3,24,0,1;2,32,0,1;4,32,0,1" --out gpurun_out/ab_128.json > gpurun_out/ab_128.log 2>&1
echo done
