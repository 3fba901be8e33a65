#!/bin/bash
# round-end measurement set on one B200 (run through gpurun): bench lines, launch lists, one full ncu capture
mkdir -p gpurun_out
T=${1:-r02}
timeout 600 python bench.py > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_c3_reference.json 2>> gpurun_out/${T}_bench_c3.err
timeout 300 python bench.py --workload C2 --no-e2e > gpurun_out/${T}_bench_c2.json 2>> gpurun_out/${T}_bench_c3.err
timeout 300 python bench.py --workload C4 --no-e2e > gpurun_out/${T}_bench_c4.json 2>> gpurun_out/${T}_bench_c3.err
timeout 300 python bench.py --workload REF --steps 10 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench_c3.err
timeout 300 python bench.py --workload REF --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref_reference.json 2>> gpurun_out/${T}_bench_c3.err
for W in C3 C2 C4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_${W}.csv python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rx_ws -s 3 -c 1 -f -o gpurun_out/${T}_prof_c3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_detect_lean -s 3 -c 1 -f -o gpurun_out/${T}_prof_c4 python bench.py --workload C4 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_ncu_c4.log 2>&1
tail -c 600 gpurun_out/${T}_bench_c3.json; echo; tail -c 400 gpurun_out/${T}_bench_c3_reference.json; echo; tail -c 300 gpurun_out/${T}_bench_ref.json; echo; tail -c 300 gpurun_out/${T}_bench_ref_reference.json; tail -3 gpurun_out/${T}_bench_c3.err
