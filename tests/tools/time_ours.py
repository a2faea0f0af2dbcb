"""GPU diagnostic (run under gpurun): per-stage times of the product alone on one of the named workloads, plus CRC32
checksums of its outputs, so two builds / two kernel variants (e.g. APD_WEAK_IMPL=old vs the default) can be compared
for speed AND bit-equality without running the reference again. Appends one JSON line to gpurun_out/time_ours.jsonl.

    python tests/tools/time_ours.py cfg3s [reps] [tag]
"""
import json, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_tools as T
from apd_mvs_b200 import engine as E

CASES = {
    "cfg2": dict(W=3111, H=2074, S=9, iters=3),
    "mid": dict(W=1024, H=768, S=9, iters=3),
    "cfg3": dict(W=6221, H=4146, S=9, iters=3, state=E.REFINE_ITER, geom=True, use_apd=True),
    "cfg3h": dict(W=3111, H=2074, S=9, iters=3, state=E.REFINE_ITER, geom=True, use_apd=True),
    "cfg3s": dict(W=1555, H=1036, S=9, iters=3, state=E.REFINE_ITER, geom=True, use_apd=True),
    "cfg4b": dict(W=1920, H=1080, S=10, iters=3, state=E.REFINE_ITER, geom=True, use_apd=True, weak_peak_radius=4),
}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3s"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tag = sys.argv[3] if len(sys.argv) > 3 else os.environ.get("APD_WEAK_IMPL", "default")
kw = dict(CASES[name])
case = T.build_case(kw.pop("W"), kw.pop("H"), kw.pop("S"), device="cuda", **kw)
names = T.stage_names(case["params"].max_iterations)
apd = T.make_product(case)
ms = None
for _ in range(reps):
    t0 = time.time(); apd.RunPatchMatch(); wall = 1e3 * (time.time() - t0)
    ms = apd.StageMs()
crc = {k: zlib.crc32(np.ascontiguousarray(v).tobytes()) for k, v in
       (("planes", apd.GetPlaneHypotheses()), ("states", apd.GetPixelStates()), ("views", apd.GetSelectedViews()), ("costs", apd.GetCosts()))}
rec = {"case": name, "tag": tag, "wall_ms": round(wall, 2), "total_ms": round(float(ms.sum()), 3), "crc": crc,
       "stage_ms": {n: round(float(m), 3) for n, m in zip(names, ms)}}
it = [m for n, m in zip(names, ms) if n.startswith("it")]
rec["iter_ms"] = round(float(sum(it)) / case["params"].max_iterations, 3)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/time_ours.jsonl", "a") as f:
    f.write(json.dumps(rec) + "\n")
print(json.dumps(rec))
apd.close()
