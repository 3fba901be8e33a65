#!/usr/bin/env python
"""Executed warp instructions per (source line, opcode) from an ncu report (cuda,sass view).
usage: ncu_inst_by_line.py report.ncu-rep [topn] [opcode-filter]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
opf = sys.argv[3] if len(sys.argv) > 3 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None; cur = None; curline = None
acc = collections.Counter(); byline = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if len(r) < 8 or r[0] == 'Function Name': continue
    if r[0] != '': curline = (cur, r[0]); continue
    if r[2] in ('...', '-'): continue
    d = dict(zip(hdr, r))
    try: n = int(d.get('Instructions Executed', '0') or 0)
    except ValueError: continue
    s = d['Source'].strip().split()
    op = s[1] if s[0].startswith('@') else s[0]
    if opf and not op.startswith(opf): continue
    acc[(curline, op)] += n; byline[curline] += n
tot = sum(acc.values())
print("total", tot)
print("-- by line")
for k, v in byline.most_common(topn): print(f"{100*v/tot:6.2f}% {v:12d} {k[0]}:{k[1]}")
print("-- by (line, opcode)")
for k, v in acc.most_common(topn): print(f"{100*v/tot:6.2f}% {v:12d} {k[0][0]}:{k[0][1]} {k[1]}")
