#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -12 gpurun_out/r02_pytest4.log | cut -c1-300
L=turdb_b200/libturdb_cuda.so,turdb_b200/libturdb_cuda.nospf.so
timeout 600 python tools/sweep.py --n 2000000 --dim 128 --metric 0 --gen clustered --genkw '{"centre_latent":16,"sigma":0.3}' --ef 128 --tunings "0,0,0,0,0;0,0,0,0,1;0,0,0,0,2" --libs $L --out gpurun_out/r02_e4_clu128.json > gpurun_out/r02_e4_clu128.log 2>&1
timeout 600 python tools/sweep.py --n 1000000 --dim 128 --metric 0 --gen sift_like --ef 128 --tunings "0,0,0,0,0;0,0,0,0,1" --libs $L --out gpurun_out/r02_e4_sift128.json > gpurun_out/r02_e4_sift128.log 2>&1
grep -h "^{\|===\|failed" gpurun_out/r02_e4_*.log | cut -c1-200
timeout 900 python bench.py --out gpurun_out/r02_bench_n1.json > gpurun_out/r02_bench_n1.log 2>&1; tail -c 3000 gpurun_out/r02_bench_n1.log
