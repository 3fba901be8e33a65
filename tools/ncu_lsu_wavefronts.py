#!/usr/bin/env python
"""Per-instruction LSU cost from an ncu report's source page: shared-memory wavefronts, global
L1 tag requests and L2 sectors, grouped by opcode and listed by instruction.
usage: ncu_lsu_wavefronts.py report.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = dict(name=line, rows=[]); blocks.append(cur); continue
    if cur is not None: cur["rows"].append(line)
for b in blocks:
    if want not in b["name"]: continue
    rd = csv.DictReader(io.StringIO("\n".join(b["rows"])))
    rows = list(rd)
    f = lambda r, k: float(r.get(k) or 0)
    tot_w = sum(f(r, "L1 Wavefronts Shared") for r in rows)
    tot_i = sum(f(r, "L1 Wavefronts Shared Ideal") for r in rows)
    tot_t = sum(f(r, "L1 Tag Requests Global") for r in rows)
    tot_s = sum(f(r, "L2 Theoretical Sectors Global") for r in rows)
    tot_inst = sum(f(r, "Instructions Executed") for r in rows)
    print(b["name"][:120])
    print(f"warp-inst {tot_inst:.3e}  shared wavefronts {tot_w:.3e} (ideal {tot_i:.3e})  global tag requests {tot_t:.3e}  L2 sectors {tot_s:.3e}")
    byop = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    for r in rows:
        op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
        a = byop[op]
        a[0] += f(r, "Instructions Executed"); a[1] += f(r, "L1 Wavefronts Shared"); a[2] += f(r, "L1 Wavefronts Shared Ideal")
        a[3] += f(r, "L1 Tag Requests Global"); a[4] += f(r, "L2 Theoretical Sectors Global")
    print("opcode                      inst     %inst   sh.wavefronts  ideal   gl.tags   L2 sectors")
    for op, a in sorted(byop.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{op:24s} {a[0]:11.3e} {100*a[0]/tot_inst:6.2f}% {a[1]:11.3e} {a[2]:11.3e} {a[3]:11.3e} {a[4]:11.3e}")
    print("top instructions by shared wavefronts:")
    for r in sorted(rows, key=lambda r: -f(r, "L1 Wavefronts Shared"))[:30]:
        print(f"  {f(r,'L1 Wavefronts Shared'):10.3e} ideal {f(r,'L1 Wavefronts Shared Ideal'):10.3e} exec {f(r,'Instructions Executed'):10.3e}  {r['Source'].strip()[:70]}")
