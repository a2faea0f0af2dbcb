"""GPU diagnostic: how often do neighbouring pixels hold bit-identical planes during a run (a duplicate candidate plane at
a pixel costs a full multi-view NCC evaluation whose result is already known)? Usage: python tests/tools/dup_stats.py cfg3s"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_tools as T
from apd_mvs_b200 import engine as E
CASES = {"cfg2q": dict(W=1555, H=1037, S=9, iters=3),
         "cfg3s": dict(W=1555, H=1036, S=9, iters=3, state=E.REFINE_ITER, geom=True, use_apd=True)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3s"
kw = dict(CASES[name]); W, H = kw["W"], kw["H"]
case = T.build_case(kw.pop("W"), kw.pop("H"), kw.pop("S"), device="cuda", **kw)
names = T.stage_names(case["params"].max_iterations)
apd = T.make_product(case)
for s, nm in enumerate(names):
    if not ("K5" in nm or "strong red" in nm or "weak red" in nm):
        continue
    apd.RunPatchMatch(stage_end=s)
    p = apd.GetPlaneHypotheses().reshape(H, W, 4).view(np.uint32)
    st = apd.GetPixelStates().reshape(H, W)
    def same(dy, dx):
        a = p[max(0, dy):H + min(0, dy), max(0, dx):W + min(0, dx)]
        b = p[max(0, -dy):H + min(0, -dy), max(0, -dx):W + min(0, -dx)]
        return (a == b).all(-1)
    # checkerboard candidates come from same-colour... no: from the OTHER colour: offsets with odd |dx|+|dy|
    near = [(0, 1), (1, 0), (0, -1), (-1, 0)]
    far = [(0, 3), (3, 0), (0, -3), (-3, 0), (1, 2), (2, 1), (-1, 2), (2, -1)]
    n1 = np.mean([same(dy, dx).mean() for dy, dx in near]); n3 = np.mean([same(dy, dx).mean() for dy, dx in far])
    # at least one of the 4 direct neighbours identical to the centre / two direct neighbours identical to each other
    c = p[1:-1, 1:-1]
    nb = [p[1:-1, 2:], p[2:, 1:-1], p[1:-1, :-2], p[:-2, 1:-1]]
    any_c = np.zeros(c.shape[:2], bool); any_pair = np.zeros(c.shape[:2], bool)
    for i in range(4):
        any_c |= (nb[i] == c).all(-1)
        for j in range(i + 1, 4):
            any_pair |= (nb[i] == nb[j]).all(-1)
    print(f"{nm:22s} identical to a given direct neighbour {n1:.3f}, to a neighbour 3 away {n3:.3f}; centre == some direct neighbour {any_c.mean():.3f}; "
          f"two direct neighbours identical {any_pair.mean():.3f}; WEAK share {(st == 1).mean():.3f}", flush=True)
apd.close()
