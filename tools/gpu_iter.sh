#!/bin/bash
# one GPU iteration: parity tests, a short bench, and (optionally) one ncu capture of the fused kernel
# usage: tools/gpu_iter.sh <tag> [ncu]
TAG=${1:-iter}
mkdir -p gpurun_out
timeout ${PYTEST_TIMEOUT:-600} python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$TAG.log 2>&1
tail -1 gpurun_out/bench_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], d['clocks'])" || tail -5 gpurun_out/bench_$TAG.log
if [ "$2" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rx_fused -s 3 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
