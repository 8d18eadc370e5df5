#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/r02_sanitize_plain.log 2>&1; echo "plain rc=$?" >> gpurun_out/r02_sanitize_plain.log
tail -6 gpurun_out/r02_sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitize_memcheck.log
grep -m 12 -A6 "Invalid\|out of bounds\|ERROR SUMMARY\|memcheck rc" gpurun_out/r02_sanitize_memcheck.log | cut -c1-200 | head -60
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -25 gpurun_out/r02_pytest3.log | cut -c1-300
L=turdb_b200/libturdb_cuda.so,turdb_b200/libturdb_cuda.c24.so,turdb_b200/libturdb_cuda.c28.so,turdb_b200/libturdb_cuda.c32.so
timeout 600 python tools/sweep.py --n 2000000 --dim 128 --metric 0 --gen clustered --genkw '{"centre_latent":16,"sigma":0.3}' --ef 128 --tunings "0,0,0,0,0;0,0,0,0,1" --libs $L --out gpurun_out/r02_e3_clu128.json > gpurun_out/r02_e3_clu128.log 2>&1
timeout 600 python tools/sweep.py --n 1000000 --dim 128 --metric 0 --gen sift_like --ef 128 --tunings "0,0,0,0,0;0,0,0,0,1" --libs $L --out gpurun_out/r02_e3_sift128.json > gpurun_out/r02_e3_sift128.log 2>&1
grep -h "^{\|===\|failed" gpurun_out/r02_e3_*.log | cut -c1-200
