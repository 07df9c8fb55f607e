#!/usr/bin/env python
"""Aggregates the per-instruction counters of an ncu source page (ncu -i X.ncu-rep --page source --csv) by code region:
loops (backward branches) and straight-line stretches between them.  Prints executed warp instructions, share, stall samples.
    python tools/ncu_regions.py source.csv [min_loop_instr]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iws, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
ins = []
for r in rows[2:]:
    if len(r) <= iex:
        continue
    ins.append((int(r[ia], 16), r[isrc].strip(), int(r[iex] or 0), int(r[isamp] or 0), int(r[iws] or 0), int(r[iwi] or 0)))
base = ins[0][0]
addr = {a - base: i for i, (a, *_rest) in enumerate(ins)}
tot = sum(x[2] for x in ins)
tots = sum(x[3] for x in ins)
minn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
loops = []
for i, (a, t, ex, sm, ws, wi) in enumerate(ins):
    m = re.search(r"BRA.*0x([0-9a-f]+)", t)
    if m:
        tg = int(m.group(1), 16)
        if tg >= base:
            tg -= base
        if tg in addr and tg < a - base and i - addr[tg] + 1 >= minn:
            loops.append((addr[tg], i))
# keep innermost / non-overlapping: sort by size, drop loops containing an accepted one
loops.sort(key=lambda l: l[1] - l[0])
acc = []
for l in loops:
    if not any(l[0] <= a[0] and a[1] <= l[1] for a in acc):
        acc.append(l)
acc.sort()
print(f"total warp instructions {tot:,}  samples {tots:,}")
pos = 0
def show(name, s, e):
    ex = sum(x[2] for x in ins[s:e + 1]); sm = sum(x[3] for x in ins[s:e + 1]); ws = sum(x[4] for x in ins[s:e + 1]); wi = sum(x[5] for x in ins[s:e + 1])
    if ex == 0 and sm == 0:
        return
    print(f"{name:28s} {hex(ins[s][0]-base):>8s}-{hex(ins[e][0]-base):<8s} n={e-s+1:5d} exec {ex:>14,} ({100*ex/tot:5.1f}%)  samples {100*sm/max(tots,1):5.1f}%  smem wavefronts {ws:,} (ideal {wi:,})")
for k, (s, e) in enumerate(acc):
    if s > pos:
        show("  straight", pos, s - 1)
    show(f"loop {k}", s, e)
    pos = e + 1
if pos < len(ins):
    show("  straight", pos, len(ins) - 1)
