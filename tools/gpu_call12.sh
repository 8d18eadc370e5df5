#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sql_operator.py -m gpu -x -q > gpurun_out/pytest_exact.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_exact.log
timeout 300 python tools/exact_probe.py --debug --dim 384 --metric 1 --gen latent --out gpurun_out/exact_probe_384.json > gpurun_out/exact_384.log 2>&1
timeout 300 python tools/exact_probe.py --debug --dim 128 --metric 0 --out gpurun_out/exact_probe_128.json > gpurun_out/exact_128.log 2>&1
timeout 800 python tools/run_configs.py --only config5_full --out gpurun_out/config5_full.json > gpurun_out/config5_full.log 2>&1
tail -3 gpurun_out/pytest_exact.log; cat gpurun_out/exact_384.log gpurun_out/exact_128.log | cut -c1-500
