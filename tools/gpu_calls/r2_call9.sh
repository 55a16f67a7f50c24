#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bf16_table or gather_rows_f32 or launches_only" > gpurun_out/r2c9_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -3 gpurun_out/r2c9_pytest.log
timeout 100 python tools/gather_bench.py 2>&1 | tail -2
timeout 200 python bench.py --workload train --steps 30 --warmup 5 2>gpurun_out/r2c9_train_err.log | tee gpurun_out/r2c9_train.json | cut -c1-700
tail -3 gpurun_out/r2c9_train_err.log
