#!/usr/bin/env python
"""Warp instructions and stall samples per CUDA source line of one kernel in an ncu report (needs -lineinfo and
--import-source on). Usage: tools/ncu_lines.py <rep> <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, agg = None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0].isdigit():
        # line summary: [line, source..., '-', '-', samples, notissued, nsamples, inst, thread inst, ...]
        try:
            k = next(i for i in range(1, len(r) - 1) if r[i] == "-" and r[i + 1] == "-")
        except StopIteration:
            continue
        try: agg.append((fname, int(r[0]), int(r[k + 5]), int(r[k + 2]), ",".join(r[1:k])[:110]))
        except (ValueError, IndexError): pass
tot = sum(a[2] for a in agg); smp = sum(a[3] for a in agg)
print(f"total warp instructions {tot}, samples {smp}")
for f, l, n, s, src in sorted(agg, key=lambda a: -a[2])[:top]:
    print(f"{100.0 * n / tot:5.1f}% inst {100.0 * s / max(smp, 1):5.1f}% smp  {f}:{l:<4d} {src.strip()}")
