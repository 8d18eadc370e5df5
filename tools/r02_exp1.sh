#!/bin/bash
# round 2, experiment 1: parity of the direct traversal form + A/B of its register-landing variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
L=turdb_b200/libturdb_cuda.so,turdb_b200/libturdb_cuda.v81.so,turdb_b200/libturdb_cuda.v161.so,turdb_b200/libturdb_cuda.nopf.so
timeout 600 python tools/sweep.py --n 1000000 --dim 128 --metric 0 --gen sift_like --ef 128 --tunings "0,0,0,0,0;0,0,0,0,1" --libs $L --out gpurun_out/r02_e1_sift128.json > gpurun_out/r02_e1_sift128.log 2>&1
timeout 600 python tools/sweep.py --n 4000000 --dim 128 --metric 0 --gen clustered --genkw '{"centre_latent":16,"sigma":0.3}' --ef 128 --tunings "0,0,0,0,0;0,0,0,0,1" --libs $L --out gpurun_out/r02_e1_clu128.json > gpurun_out/r02_e1_clu128.log 2>&1
timeout 600 python tools/sweep.py --n 1000000 --dim 384 --metric 1 --ef 128 --tunings "0,0,0,0,0;0,0,0,0,2" --libs turdb_b200/libturdb_cuda.so,turdb_b200/libturdb_cuda.v161.so --out gpurun_out/r02_e1_lat384.json > gpurun_out/r02_e1_lat384.log 2>&1
grep -h "^{\|===\|failed" gpurun_out/r02_e1_*.log | cut -c1-260
