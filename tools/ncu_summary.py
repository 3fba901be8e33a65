#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hottest source lines (by stall samples)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
d = dict(zip(hdr, zip(units, vals)))
for k in keys:
    if k in d: print(f"{k:85s} {d[k][1]:>16s} {d[k][0]}")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(d[h][1])
        if v > 0.1: print(f"  stall {h[34:-28]:28s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}; tot = 0; tots = 0
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''): continue
    try: ln = int(r[0]); inst = int(r[7]); samp = int(r[4])
    except ValueError: continue
    agg[(cur, ln)] = (inst, samp, r[1].strip()[:90]); tot += inst; tots += samp
print("total warp-inst", tot, "samples", tots)
pf = collections.Counter(); ps = collections.Counter()
for (f, l), (i, s, _) in agg.items(): pf[f] += i; ps[f] += s
for f in pf: print(f"  {f:32s} inst {pf[f]/tot*100:5.1f}%  samples {ps[f]/tots*100:5.1f}%")
for (f, l), (i, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f[:22]:22s}:{l:4d} inst {i/tot*100:5.2f}% smp {s/tots*100:5.2f}%  {t}")
