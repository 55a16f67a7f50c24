#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "maxpool or meanpool" > gpurun_out/r2c2_pytest.log 2>&1; echo "[pytest k4] rc=$?"; tail -5 gpurun_out/r2c2_pytest.log
TC_CHECK_SKIP_TESTS=1 timeout 300 python tools/tc_check.py > gpurun_out/r2c2_tc.log 2>&1; echo "[tc_check] rc=$?"; tail -12 gpurun_out/r2c2_tc.log
timeout 300 python bench.py --aggregator maxpool --steps 50 --warmup 10 --cpu-batches 0 > gpurun_out/r2c2_bench_maxpool.log 2>&1; echo "[bench maxpool] rc=$?"; tail -1 gpurun_out/r2c2_bench_maxpool.log | cut -c1-600
