#!/bin/bash
# exact path: where the time goes — launch list, TMEM-read / tensor-pipe contention diagnostics, slice growth
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SH="--dim 384 --metric 1 --gen gaussian_latent"
for G in 4 8 13 16; do
  echo "== growth $G"; TURDB_EXACT_GROWTH=$G timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact4_growth$G.json 2>&1 | tail -1 | cut -c180-330
done
echo "== auto growth, pair"; TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py $SH --out gpurun_out/r02_exact4_auto_pair.json 2>&1 | tail -1 | cut -c180-330
for D in 1 2; do for P in 0 1; do
  echo "== DIAG $D pair $P (results wrong by design)"; TURDB_EXACT_GROWTH=4 TURDB_EXACT_DIAG=$D TURDB_EXACT_PAIR=$P timeout 200 python tools/exact_probe.py $SH --debug --out gpurun_out/r02_exact4_diag${D}_pair$P.json 2>&1 | tail -2 | cut -c1-330
done; done
echo "== 128-d and 768-d, auto growth"
timeout 200 python tools/exact_probe.py --dim 128 --metric 0 --gen sift_like --out gpurun_out/r02_exact4_auto_128.json 2>&1 | tail -1 | cut -c180-330
TURDB_EXACT_PAIR=1 timeout 200 python tools/exact_probe.py --dim 768 --metric 2 --gen gaussian_latent --out gpurun_out/r02_exact4_auto_768_pair.json 2>&1 | tail -1 | cut -c180-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'exact_|to_half|query_slack|col_bias|fill_' -c 200 --csv --log-file gpurun_out/r02_launches_exact_probe.csv \
   python tools/exact_probe.py $SH --reps 1 --out gpurun_out/ncu_dummy.json > gpurun_out/r02_ncu_exact_list.log 2>&1; tail -1 gpurun_out/r02_ncu_exact_list.log | cut -c1-200
