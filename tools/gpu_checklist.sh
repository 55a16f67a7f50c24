#!/usr/bin/env bash
# Developer checklist for a GPU box (run from the repo root, each stage under its own timeout so a hang cannot eat the
# box): full parity suite, the tensor-core checks incl. the opt-in TMA gather4 producer of K4, the three bench arms and
# the launch list.  Outputs go to gpurun_out/ (scratch); copy what should be judged into profiles/.
#   usage: tools/gpu_checklist.sh [stage ...]     stages: tests tc g4 bench maxpool launches   (default: all)
set -u
mkdir -p gpurun_out
stages=("$@")
[ ${#stages[@]} -eq 0 ] && stages=(tests tc g4 bench maxpool launches)
for st in "${stages[@]}"; do
  case "$st" in
    tests)    timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ;;
    tc)       timeout 200 python tools/tc_check.py 2>&1 | tail -12 ;;
    g4)       TC_CHECK_G4=1 timeout 240 python tools/tc_check.py 2>&1 | tail -8 ;;
    bench)    timeout 400 python bench.py --steps 500 --warmup 30 2>&1 | tail -1 | tee gpurun_out/bench_mean.json ;;
    maxpool)  timeout 400 python bench.py --aggregator maxpool --math bf16 --steps 100 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_maxpool.json ;;
    launches) timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv \
                --log-file gpurun_out/launches.csv python bench.py --steps 60 --warmup 30 --cpu-batches 0 > gpurun_out/launches.log 2>&1
              tail -3 gpurun_out/launches.log ;;
    *) echo "unknown stage $st" ;;
  esac
  echo "[stage $st] rc=$?"
done
