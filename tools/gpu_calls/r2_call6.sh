#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "maxpool or meanpool or rmat or csr_sampled" > gpurun_out/r2c6_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -4 gpurun_out/r2c6_pytest.log
timeout 200 python tools/k4_matrix.py > gpurun_out/r2c6_matrix.log 2>&1; echo "[matrix] rc=$?"; tail -9 gpurun_out/r2c6_matrix.log
timeout 100 python bench.py --workload rmat --rmat-scale 25 --steps 20 --warmup 5 --repeats 3 > gpurun_out/r2c6_rmat25.log 2>&1; echo "[rmat25] rc=$?"; tail -1 gpurun_out/r2c6_rmat25.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['graph'], d['value'], d['gather_kernel_ms'])"
