#!/usr/bin/env python
"""Per-role register use of a setmaxnreg kernel: splits the SASS of one kernel at the USETMAXREG
instructions and reports instruction count, highest register index and spill instructions per region.
usage: sass_regions.py <object-or-so> <kernel-substring>"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
blocks = out.split("\t\tFunction : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if sys.argv[2] not in name: continue
    lines = [l for l in b.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
    marks = [i for i, l in enumerate(lines) if "SETMAXREG" in l]
    print(name, len(lines), "instructions")
    bounds = [0] + marks + [len(lines)]
    for a, e in zip(bounds[:-1], bounds[1:]):
        mx = 0; sp = 0; ops = {}
        for l in lines[a:e]:
            for m in re.finditer(r"\bR(\d+)\b", l): mx = max(mx, int(m.group(1)))
            if re.search(r"\b(STL|LDL)\b", l): sp += 1
        print(f"  [{a:6d},{e:6d}) {lines[a].split(';')[0].split('*/')[1].strip()[:60]:60s} maxR {mx:3d} spill-insts {sp}")
