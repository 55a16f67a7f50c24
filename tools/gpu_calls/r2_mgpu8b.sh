#!/usr/bin/env bash
set -u
N=8
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload rmat --rmat-scale 27 --rmat-nodes 100000000 --steps 50 --warmup 10 > gpurun_out/r2m8_rmat.log 2>&1; echo "[rmat full] rc=$?"; tail -1 gpurun_out/r2m8_rmat.log | cut -c1-2400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload unsup --steps 20 --warmup 5 > gpurun_out/r2m8_unsup.log 2>&1; echo "[unsup 8] rc=$?"; tail -1 gpurun_out/r2m8_unsup.log | cut -c1-1200
GS_HALO_CACHE_SWEEP=0.25 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/r2m8_bench.log 2>&1; echo "[bench 8] rc=$?"; tail -1 gpurun_out/r2m8_bench.log | cut -c1-300
grep -i "error\|Traceback" -A8 gpurun_out/r2m8_rmat.log gpurun_out/r2m8_unsup.log gpurun_out/r2m8_bench.log | head -40
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3
