#!/bin/bash
# why are the early slices 5x slower per tile at 128-d?  ncu --set full of pass 4 (72 tiles) and per-launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SH="--dim 128 --metric 2 --gen gaussian_latent"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_gemm_filter -s 11 -c 1 -o gpurun_out/r02_prof_exact_128_pass4 \
    python tools/exact_probe.py $SH --reps 1 --out gpurun_out/ncu_dummy.json > gpurun_out/r02_ncu_exact_128_pass4.log 2>&1; tail -1 gpurun_out/r02_ncu_exact_128_pass4.log | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'exact_|to_half|query_slack' -c 200 --csv --log-file gpurun_out/r02_launches_exact_128_ip_final.csv \
   python tools/exact_probe.py $SH --reps 1 --out gpurun_out/ncu_dummy.json > /dev/null 2>&1; echo list done
