#!/usr/bin/env python
"""Stall samples of a warp-specialised kernel split by role: the SASS listing of the ncu source page is
cut at the USETMAXREG / EXIT instructions that delimit the producer, FFT and detect code regions.
usage: ncu_roles.py report.ncu-rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None; insts = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): insts.append(dict(zip(hdr, r)))
print(len(insts), "SASS instructions")
def num(d, k):
    try: return float(d.get(k) or 0)
    except ValueError: return 0.0
cuts = [i for i, d in enumerate(insts) if 'SETMAXREG' in d['Source'] or d['Source'].strip().endswith('EXIT ;') or ' EXIT' in d['Source']]
bounds = sorted(set([0] + cuts + [len(insts)]))
tot = sum(num(d, '# Samples') for d in insts)
stall_keys = [k for k in hdr if k.startswith('stall_') and '(' not in k]
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = insts[a:b]
    s = sum(num(d, '# Samples') for d in seg)
    if s < 0.002 * tot: continue
    ex = sum(num(d, 'Instructions Executed') for d in seg)
    st = collections.Counter()
    for d in seg:
        for k in stall_keys: st[k[6:]] += num(d, k)
    top = " ".join(f"{k}={v/s*100:.0f}%" for k, v in st.most_common(6))
    print(f"[{a:6d},{b:6d}) {insts[a]['Source'].strip()[:44]:44s} samples {s/tot*100:5.1f}%  warp-inst {ex:.3e}  {top}")
