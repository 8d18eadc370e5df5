#!/bin/bash
# why is the L2 (bias) instantiation 1.5x slower than IP at 128-d?  slack-band diagnostic + per-launch lists
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A="--dim 128 --metric 0 --gen gaussian_latent"; B="--dim 128 --metric 2 --gen gaussian_latent"
echo "== L2 slack off (uncertified diagnostic)"; TURDB_EXACT_SLACK_SCALE=0 timeout 200 python tools/exact_probe.py $A --out gpurun_out/r02_exact9_l2_noslack.json 2>&1 | tail -1 | cut -c1-20,180-420
echo "== L2 bf16"; TURDB_EXACT_FORCE_BF16=1 timeout 200 python tools/exact_probe.py $A --out gpurun_out/r02_exact9_l2_bf16.json 2>&1 | tail -1 | cut -c1-20,180-420
K='regex:exact_|to_half|query_slack|col_bias|fill_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r02_launches_exact_128_l2.csv python tools/exact_probe.py $A --reps 1 --out gpurun_out/ncu_dummy.json > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r02_launches_exact_128_ip.csv python tools/exact_probe.py $B --reps 1 --out gpurun_out/ncu_dummy.json > /dev/null 2>&1
echo done
