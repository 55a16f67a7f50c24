#!/usr/bin/env bash
# round-2 GPU call 1 (1 GPU): parity suite, K4 producer variants, short bench (mean + config-3 pass), K4 ncu on real ids
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c1_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -5 gpurun_out/r2c1_pytest.log
TC_CHECK_SKIP_TESTS=1 TC_CHECK_G4=1 timeout 300 python tools/tc_check.py > gpurun_out/r2c1_tc.log 2>&1; echo "[tc_check] rc=$?"; tail -14 gpurun_out/r2c1_tc.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/r2c1_bench.log 2>&1; echo "[bench] rc=$?"; tail -1 gpurun_out/r2c1_bench.log | cut -c1-1500
timeout 300 ncu --set full --clock-control none --import-source on -k regex:maxpool_mlp -s 4 -c 1 -o gpurun_out/r2c1_k4 python tools/maxpool_bench.py > gpurun_out/r2c1_ncu.log 2>&1; echo "[ncu k4] rc=$?"; tail -3 gpurun_out/r2c1_ncu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
