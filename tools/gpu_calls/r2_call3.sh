#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "maxpool or meanpool" > gpurun_out/r2c3_pytest.log 2>&1; echo "[pytest k4] rc=$?"; tail -5 gpurun_out/r2c3_pytest.log
timeout 200 python tools/k4_matrix.py > gpurun_out/r2c3_matrix.log 2>&1; echo "[matrix] rc=$?"; tail -14 gpurun_out/r2c3_matrix.log
GS_TUNING=k4_kernel=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:maxpool_mlp -s 4 -c 1 -o gpurun_out/r2c3_k4wide python tools/maxpool_bench.py > gpurun_out/r2c3_ncu.log 2>&1; echo "[ncu k4] rc=$?"; tail -2 gpurun_out/r2c3_ncu.log
