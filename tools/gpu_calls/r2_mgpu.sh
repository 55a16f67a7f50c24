#!/usr/bin/env bash
# multi-GPU call: N from $1 (default 2).  2-GPU parity test of the partitioned path, then the bench at N GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 400 python -m pytest tests/test_parallel.py -x -q -m gpu > gpurun_out/r2m_pytest_$N.log 2>&1; echo "[pytest parallel] rc=$?"; tail -4 gpurun_out/r2m_pytest_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-100} --warmup 10 > gpurun_out/r2m_bench_$N.log 2>&1; echo "[bench $N] rc=$?"; tail -1 gpurun_out/r2m_bench_$N.log | cut -c1-3000
grep -i "error\|Traceback" -A5 gpurun_out/r2m_bench_$N.log | head -30
