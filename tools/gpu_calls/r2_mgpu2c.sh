#!/usr/bin/env bash
set -u
N=2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parallel.py -x -q -m gpu -k "two_gpus" > gpurun_out/r2m3_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -4 gpurun_out/r2m3_pytest.log
for st in 1 0; do
GS_HALO_STAGING=$st GS_HALO_CACHE_ROWS=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$st bench.py --gpus $N --steps 50 --warmup 10 --repeats 8 > gpurun_out/r2m3_bench_st$st.log 2>&1; echo "[bench staging=$st] rc=$?"
tail -1 gpurun_out/r2m3_bench_st$st.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); p=d['partitioned']
print('value',d['value'],'ms/step',d['ms_per_step'],'gather_ms',p['gather_kernel_ms'],'remote',p['remote_row_fraction_after_replicas'],'nvlink',p['nvlink_GBps_per_gpu'], 'uniq', p.get('unique_remote_rows_per_step'), 'replicated', d['replicated']['value'])
"
done
grep -i "error\|Traceback" -A8 gpurun_out/r2m3_bench_st1.log gpurun_out/r2m3_bench_st0.log | head -30
