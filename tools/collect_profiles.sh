#!/bin/bash
# copies the outputs of tools/final_measure.sh <tag> from gpurun_out/ into profiles/ (r02_*) and stamps
# profiles/traffic.json with the CURRENT kernel-source hash: run it on the commit the capture was taken from
T=${1:?tag}; P=profiles
for f in c3 c3_reference c2 c4 ref ref_reference; do cp gpurun_out/${T}_bench_$f.json $P/r02_bench_$f.json; done
for W in C3 C2 C4; do grep '^"' gpurun_out/${T}_launches_$W.csv > $P/r02_launches_$W.csv; done
python tools/ncu_traffic.py gpurun_out/${T}_prof_c3.ncu-rep C3 > /dev/null
python tools/ncu_summary.py gpurun_out/${T}_prof_c3.ncu-rep > $P/r02_ncu_c3_summary.txt 2>&1
python tools/ncu_roles.py gpurun_out/${T}_prof_c3.ncu-rep > $P/r02_ncu_c3_roles.txt 2>&1
python tools/ncu_sass_stalls.py gpurun_out/${T}_prof_c3.ncu-rep > $P/r02_ncu_c3_sass_stalls.txt 2>&1
python tools/ncu_summary.py gpurun_out/${T}_prof_c4.ncu-rep > $P/r02_ncu_c4_detect_summary.txt 2>&1
cp gpurun_out/f64_C3.json $P/r02_f64_C3.json; cp gpurun_out/f64_C4.json $P/r02_f64_C4.json
