#!/bin/bash
# A/B of several builds: tools/ab_libs.sh lib_a.so lib_b.so ...   (same box, same process order; OUT=outputs)
mkdir -p gpurun_out
for L in "$@" "$1"; do
  RUB_MIMO_LIB=$PWD/rub_mimo_b200/$L python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --outputs ${OUT:-eq+llr+bits} ${EXTRA} > gpurun_out/ab_tmp.log 2>&1
  tail -1 gpurun_out/ab_tmp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']), d['clocks']['sm_mhz'])" || tail -5 gpurun_out/ab_tmp.log
done
