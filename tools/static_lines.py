#!/usr/bin/env python
"""Static code size of one kernel by source line (no GPU needed): SASS instructions per `//## File ..., line N` run of
`nvdisasm --print-line-info`. Usage: tools/static_lines.py <object.o> <kernel name substring> [top N]
The instruction cache is a measured limiter of these kernels (L0 ~6 KB, L1.5 32 KB): cold code that the compiler unrolled
is footprint the hot loops pay for."""
import collections, os, re, subprocess, sys, tempfile
obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cnt, src, cur, infn, total = collections.Counter(), None, None, False, 0
for l in txt.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l) or re.match(r"\.section\s+\.text\.(\S+),", l)
    if m:
        infn = pat in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and cur:
        cnt[cur] += 1; total += 1
print(f"{total} SASS instructions in kernels matching '{pat}'")
for (f, n), c in cnt.most_common(top):
    print(f"{c:6d}  {f}:{n}")
