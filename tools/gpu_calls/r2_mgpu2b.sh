#!/usr/bin/env bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_parallel.py tests/test_gpu_parity.py -x -q -m gpu -k "two_gpus or csr_sampled" > gpurun_out/r2m2_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -4 gpurun_out/r2m2_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload unsup --steps 20 --warmup 5 > gpurun_out/r2m2_unsup_$N.log 2>&1; echo "[unsup $N] rc=$?"; tail -1 gpurun_out/r2m2_unsup_$N.log | cut -c1-1200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload rmat --rmat-scale ${RS:-22} --rmat-nodes ${RN:-0} --steps 50 --warmup 10 > gpurun_out/r2m2_rmat_$N.log 2>&1; echo "[rmat $N] rc=$?"; tail -1 gpurun_out/r2m2_rmat_$N.log | cut -c1-2200
grep -i "error\|Traceback" -A8 gpurun_out/r2m2_unsup_$N.log gpurun_out/r2m2_rmat_$N.log | head -40
