#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A="2000000 --dim 128 --metric 0 --gen clustered --genkw {\"centre_latent\":16,\"sigma\":0.3} --ef 128"
TURDB_CUDA_LIB=$PWD/turdb_b200/libturdb_cuda.v81d.so timeout 600 python tools/ncu_target.py $A --debug > gpurun_out/r02_e2_dbg.log 2>&1
TURDB_CUDA_LIB=$PWD/turdb_b200/libturdb_cuda.v81.so timeout 600 python tools/ncu_target.py $A > gpurun_out/r02_e2_plain.log 2>&1 &&
TURDB_CUDA_LIB=$PWD/turdb_b200/libturdb_cuda.v81.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_warp_kernel -s 6 -c 1 -o gpurun_out/r02_prof_direct_clu128 \
    python tools/ncu_target.py $A > gpurun_out/r02_e2_ncu.log 2>&1
tail -4 gpurun_out/r02_e2_dbg.log gpurun_out/r02_e2_plain.log; tail -3 gpurun_out/r02_e2_ncu.log
