#!/usr/bin/env python
"""Top stalled SASS instructions inside an index range of the kernel's SASS listing (see ncu_roles.py for the ranges).
usage: ncu_region_top.py report.ncu-rep first last [topn]"""
import csv, subprocess, sys
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]); topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None; insts = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): insts.append(dict(zip(hdr, r)))
def num(d, k):
    try: return float(d.get(k) or 0)
    except ValueError: return 0.0
seg = insts[a:b]
tot = sum(num(d, '# Samples') for d in seg)
sk = [k for k in hdr if k.startswith('stall_') and '(' not in k]
for i, d in sorted(enumerate(seg), key=lambda t: -num(t[1], '# Samples'))[:topn]:
    rs = sorted(((k[6:], num(d, k)) for k in sk), key=lambda kv: -kv[1])[:3]
    print(f"{a+i:6d} {num(d,'# Samples')/tot*100:5.2f}% exec {num(d,'Instructions Executed'):.2e} {d['Source'].strip()[:60]:60s} " + " ".join(f"{k}={v:.0f}" for k, v in rs if v > 0))
