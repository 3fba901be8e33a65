#!/usr/bin/env python
"""Top SASS instructions by stall samples with their dominant stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None; cur = None; curline = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if len(r) < 8 or r[0] == 'Function Name': continue
    if r[0] != '': curline = (cur, r[0]); continue
    if r[2] in ('...', '-'): continue
    d = dict(zip(hdr, r))
    try: tot = int(d.get('# Samples', '0') or 0)
    except ValueError: continue
    reasons = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and '(' not in k and v not in ('', '-') and int(v) > 0}
    out.append((tot, d['Source'].strip()[:58], curline, reasons))
tot_all = sum(o[0] for o in out)
for o in sorted(out, key=lambda o: -o[0])[:topn]:
    rs = sorted(o[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"{o[0]/tot_all*100:5.2f}% {o[1]:58s} {o[2][0][:14]}:{o[2][1]:4s} " + " ".join(f"{k}={v}" for k, v in rs))
