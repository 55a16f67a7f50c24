#!/usr/bin/env bash
# final validation on a 2-GPU box: the whole GPU suite (incl. the 2-GPU tests), smoke, the default bench at N = 1 and N = 2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "[pytest] rc=$?"; tail -3 gpurun_out/r2g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 100 --warmup 10 --cpu-batches 2 > gpurun_out/r2g_bench1.log 2>&1; echo "[bench N=1] rc=$?"; tail -1 gpurun_out/r2g_bench1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=1 value',d['value'],'us/step',d['ms_per_step']*1e3,'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],'tensor',d['roofline_tensor']['frac'], d['roofline_tensor']['avg_kernel_ms'],'c3',d['config3']['value'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2g_bench2.log 2>&1; echo "[bench N=2] rc=$?"; tail -1 gpurun_out/r2g_bench2.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); p=d['partitioned']; print('N=2 value',d['value'],'us/step',d['ms_per_step']*1e3,'e2e',d['e2e']['value'],'remote',p['remote_row_fraction_after_replicas'],'replicas',p['replica_fraction_of_table'],'replicated',d['replicated']['value'])"
grep -i "error\|Traceback" -A6 gpurun_out/r2g_bench1.log gpurun_out/r2g_bench2.log | head -20
