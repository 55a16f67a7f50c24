#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python tools/k4_matrix.py > gpurun_out/r2c4_matrix.log 2>&1; echo "[matrix] rc=$?"; tail -12 gpurun_out/r2c4_matrix.log
timeout 100 python tools/mma_rate.py > gpurun_out/r2c4_mmarate.log 2>&1; echo "[mma_rate] rc=$?"; tail -12 gpurun_out/r2c4_mmarate.log
