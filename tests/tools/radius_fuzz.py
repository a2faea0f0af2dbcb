"""GPU diagnostic: parity sweep over weak_peak_radius values the randomised sweep does not draw (0, 1, 3, 10, 20, 28),
deformation on and off. Usage: python tests/tools/radius_fuzz.py [cases per radius]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_tools as T
from apd_mvs_b200 import engine as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rng = np.random.default_rng(555)
bad = tot = 0
for rad in (0, 1, 3, 10, 20, 28):
    for i in range(n):
        W, H, S, kw, seed = T.random_config(rng)
        kw["weak_peak_radius"] = rad
        case = T.build_case(W, H, S, device="cuda", **kw)
        d = T.final_diff(case, seed)
        ok = all(v == 0.0 for v in d.values()); bad += not ok; tot += 1
        print(f"radius {rad:2d} {W}x{H} S={S} state={kw['state']} apd={int(kw['use_apd'])} geom={int(kw['geom'])} -> {'OK' if ok else d}", flush=True)
print(f"{tot - bad} / {tot} configurations bit-identical")
sys.exit(1 if bad else 0)
