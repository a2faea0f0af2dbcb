#!/usr/bin/env python
"""Basic-block view of one kernel in an ncu report: runs of consecutive SASS instructions with the same execution
count, with their share of all warp instructions. Usage: tools/ncu_blocks.py <rep> <kernel regex> [min share %]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
h = rows[hi]; s, ie, sm = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
ins = [(r[s], int(r[ie]), int(r[sm] or 0)) for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(i[1] for i in ins); smp = sum(i[2] for i in ins)
blocks, cur = [], None
for idx, (src, n, st) in enumerate(ins):
    if cur and cur[2] == n: cur[1] += 1; cur[3] += st; cur[4] += ("TEX" in src)
    else:
        cur = [idx, 1, n, st, int("TEX" in src), src]; blocks.append(cur)
print(f"total {tot} warp instructions, {len(ins)} SASS instructions")
for b in blocks:
    share = 100.0 * b[1] * b[2] / tot
    if share >= minp: print(f"  @{b[0]:5d} len {b[1]:4d} x{b[2]:>12d} = {share:5.1f}% inst {100.0 * b[3] / max(smp, 1):5.1f}% smp  TEX {b[4]:3d}  {b[5][:60]}")
