#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_parallel.py -x -q -m gpu -k "image_layer or khop_golden or full_size_forward or graphed or single_rank or aggregators_golden" > gpurun_out/r2c8_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -4 gpurun_out/r2c8_pytest.log
for d in 3 4; do timeout 100 python bench.py --steps 200 --warmup 20 --cpu-batches 0 --no-config3 --repeats 5 --depth $d 2>gpurun_out/r2c8_bench_err_$d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('depth', d['impl_detail']['pipeline_depth'], 'value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], 'gather ms', d['kernel_ms'], 'launches/step', d['gpu_launches']/d['steps'])"; done
tail -5 gpurun_out/r2c8_bench_err_3.log
