#!/usr/bin/env bash
# final 1-GPU evidence run: parity suite, smoke, bench lines, ncu launch lists and full captures of the two dominant kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -3 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "[smoke] rc=$?"; tail -3 gpurun_out/r2f_smoke.log
timeout 400 python bench.py --steps 500 --warmup 30 > gpurun_out/r2f_bench_mean.log 2>&1; echo "[bench mean] rc=$?"; tail -1 gpurun_out/r2f_bench_mean.log | cut -c1-200
timeout 300 python bench.py --aggregator maxpool --steps 100 --warmup 10 --cpu-batches 0 > gpurun_out/r2f_bench_maxpool.log 2>&1; echo "[bench maxpool] rc=$?"; tail -1 gpurun_out/r2f_bench_maxpool.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2f_launches_mean.csv python bench.py --steps 40 --warmup 20 --cpu-batches 0 --no-config3 --repeats 1 > gpurun_out/r2f_launches_mean.log 2>&1; echo "[launches mean] rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2f_launches_maxpool.csv python bench.py --aggregator maxpool --steps 20 --warmup 10 --cpu-batches 0 --repeats 1 > gpurun_out/r2f_launches_maxpool.log 2>&1; echo "[launches maxpool] rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:maxpool_mlp -s 4 -c 1 -o gpurun_out/r2f_k4 python tools/maxpool_bench.py > gpurun_out/r2f_ncu_k4.log 2>&1; echo "[ncu k4] rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_mean -s 40 -c 1 -o gpurun_out/r2f_gather python bench.py --steps 20 --warmup 10 --cpu-batches 0 --no-config3 --repeats 1 > gpurun_out/r2f_ncu_gather.log 2>&1; echo "[ncu gather] rc=$?"
