#!/usr/bin/env bash
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-100} --warmup 10 > gpurun_out/r2m_bench_$N.log 2>&1; echo "[bench $N] rc=$?"; tail -1 gpurun_out/r2m_bench_$N.log | cut -c1-400
grep -i "error\|Traceback" -A5 gpurun_out/r2m_bench_$N.log | head -30
