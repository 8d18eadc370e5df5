#!/bin/bash
# TMEM load micro-probe (does tcgen05.ld cost tensor-pipe time? which shape?), exact path with kprime = k and growth 2/3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 tools/micro/tmem_ld_probe.bin 2>&1 | tee gpurun_out/r02_tmem_ld_probe.txt
SH="--dim 384 --metric 1 --gen gaussian_latent"
for R in 1 2; do for G in 2 3; do
  echo "== rerank $R growth $G"; TURDB_EXACT_GROWTH=$G timeout 200 python tools/exact_probe.py $SH --rerank $R --out gpurun_out/r02_exact6_r${R}_g$G.json 2>&1 | tail -1 | cut -c180-330
done; done
echo "== 128-d, 768-d rerank 1 growth 2"
TURDB_EXACT_GROWTH=2 timeout 200 python tools/exact_probe.py --dim 128 --metric 0 --gen sift_like --rerank 1 --out gpurun_out/r02_exact6_128.json 2>&1 | tail -1 | cut -c180-330
TURDB_EXACT_GROWTH=2 timeout 200 python tools/exact_probe.py --dim 768 --metric 2 --gen gaussian_latent --rerank 1 --out gpurun_out/r02_exact6_768.json 2>&1 | tail -1 | cut -c180-330
