#!/bin/bash
# DRAM traffic and duration of the dominant kernel for several builds: tools/dev/dram_ab.sh lib_a.so lib_b.so ...
mkdir -p gpurun_out
for L in "$@"; do
  RUB_MIMO_LIB=$PWD/rub_mimo_b200/$L timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_rx_ws -s 3 -c 2 --csv --log-file gpurun_out/dram_tmp.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
  python - "$L" <<'PY'
import csv,sys
rows=[r for r in csv.reader(l for l in open('gpurun_out/dram_tmp.csv') if l.startswith('"'))]
h=rows[0]; mi=h.index("Metric Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
out={}
for r in rows[1:]: out.setdefault(r[mi],[]).append((r[vi],r[ui]))
print(sys.argv[1], {k:v for k,v in out.items()})
PY
done
