#!/usr/bin/env python
"""Randomised parity sweep (run under gpurun): N random configurations of size, number of source views, run state,
use_APD, geometric term, iterations, rotate_time, top_k, weak_peak_radius and seed; every final output of the product is
compared bit for bit with the reference oracle (oracle/_ref/libapd_ref.so). Writes gpurun_out/parity_fuzz.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_tools as T
from apd_mvs_b200 import engine as E

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
rows, bad = [], 0
for i in range(n_cases):
    W, H, S, kw, curand_seed = T.random_config(rng)
    state, use_apd, geom = kw["state"], kw["use_apd"], kw["geom"]
    case = T.build_case(W, H, S, device="cuda", **kw)
    d = T.final_diff(case, curand_seed)
    ok = all(v == 0.0 for v in d.values())
    bad += 0 if ok else 1
    rows.append({"W": W, "H": H, "S": S, **{k: v for k, v in kw.items()}, "curand_seed": curand_seed, "diff": d, "ok": ok})
    print(f"[{i:3d}] {W}x{H} S={S} state={state} apd={int(use_apd)} geom={int(geom)} it={kw['iters']} rot={kw['rotate_time']} topk={kw['top_k']} -> {'OK' if ok else d}", flush=True)
print(f"{n_cases - bad} / {n_cases} configurations bit-identical")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"cases": rows, "bit_identical": n_cases - bad, "total": n_cases}, open("gpurun_out/parity_fuzz.json", "w"), indent=1)
sys.exit(1 if bad else 0)
