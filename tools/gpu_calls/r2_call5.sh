#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_parallel.py -x -q -m gpu -k "rmat or csr_sampled or single_rank or sharded" > gpurun_out/r2c5_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -4 gpurun_out/r2c5_pytest.log
timeout 300 python bench.py --workload rmat --rmat-scale 22 --steps 50 --warmup 10 > gpurun_out/r2c5_rmat.log 2>&1; echo "[rmat 1gpu] rc=$?"; tail -1 gpurun_out/r2c5_rmat.log | cut -c1-1800
timeout 300 python bench.py --workload unsup --steps 20 --warmup 5 > gpurun_out/r2c5_unsup.log 2>&1; echo "[unsup 1gpu] rc=$?"; tail -1 gpurun_out/r2c5_unsup.log | cut -c1-1500
grep -i "error\|Traceback" -A8 gpurun_out/r2c5_rmat.log gpurun_out/r2c5_unsup.log | head -40
