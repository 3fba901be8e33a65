#!/bin/bash
# development A/B of the warp-specialised fused kernel against the monolithic one (C3), under timeouts
mkdir -p gpurun_out
TAG=${1:-ws}
echo "== parity (WS)"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -x -q 2>&1 | tail -15
for v in 0 1; do
  RUB_FUSED_WS=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_ws$v.log 2>&1
  tail -1 gpurun_out/bench_${TAG}_ws$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WS=$v ms/step %.4f kernel_ms %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']), d['clocks'])" || tail -5 gpurun_out/bench_${TAG}_ws$v.log
done
if [ "$2" = "ncu" ]; then
  RUB_FUSED_WS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rx_ws -s 3 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
