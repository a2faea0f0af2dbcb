"""Tiny run of every kernel for `compute-sanitizer --tool initcheck` (slow tool: keep the case small)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import parity_tools as T
from apd_mvs_b200 import engine as E
for kw in (dict(state=E.FIRST_INIT), dict(state=E.REFINE_ITER, use_apd=True, geom=True, rotate_time=2)):
    case = T.build_case(72, 56, 2, iters=1, **kw)
    apd = T.make_product(case)
    apd.RunPatchMatch()
    print("case", kw, float(np.nanmean(T.product_state(apd)["planes"][..., 3])))
    apd.close()
from apd_mvs_b200 import pipeline as P, fusion as F
from apd_mvs_b200.scene import make_scene
import fusion_tools as FT
sc = make_scene(1002, 32, 2)
s = P.Scene(sc["images"].numpy(), sc["cameras"], P.ring_pairs(3, 2))
s.Run()
print("scene", s.ComputeRoundNum(), float(s.Depth(0).mean()))
fu = F.Fusion(3, 1002, 32)
bgr = FT.colour_images(sc["images"].numpy())
for v in range(3):
    fu.SetView(v, bgr[v], sc["cameras"][v], s.Depth(v), s.Normal(v), s.States(v))
for r, ss in P.ring_pairs(3, 2):
    fu.AddProblem(r, ss)
xyz, col = fu.RunFusion()
print("fusion", len(xyz))
fu.close(); s.close()
