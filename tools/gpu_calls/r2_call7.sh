#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "maxpool or meanpool or gather_rows_f32 or graphed or pipelined" > gpurun_out/r2c7_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -3 gpurun_out/r2c7_pytest.log
K4_MATRIX_ONLY="tmem 128 + cluster" timeout 100 python tools/k4_matrix.py > gpurun_out/r2c7_matrix.log 2>&1; echo "[matrix] rc=$?"; tail -3 gpurun_out/r2c7_matrix.log
for d in 2 4 5; do timeout 100 python bench.py --steps 200 --warmup 20 --cpu-batches 0 --no-config3 --repeats 5 --depth $d 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('depth', d['impl_detail']['pipeline_depth'], 'value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'])"; done
timeout 200 python bench.py --aggregator maxpool --steps 100 --warmup 10 --cpu-batches 0 --repeats 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('maxpool value', d['value'], 'us/step', d['ms_per_step']*1e3, 'k4 ms', d['roofline']['avg_kernel_ms'], 'frac', d['roofline']['frac'], 'launches/step', d['gpu_launches']/d['steps'])"
