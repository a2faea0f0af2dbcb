#!/bin/bash
# compute-sanitizer passes over small cases of every kernel (memcheck, racecheck, initcheck on the engine's own buffers).
# Output: gpurun_out/sanitize_*.log ; summarised in profiles/README.md.
set -u
mkdir -p gpurun_out
export PYTHONPATH=.:tests
cat > /tmp/san_case.py <<'PY'
import numpy as np, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity_tools as T
from apd_mvs_b200 import engine as E, pipeline as P
from apd_mvs_b200.scene import make_scene
for kw in (dict(state=E.FIRST_INIT), dict(state=E.REFINE_ITER, use_apd=True, geom=True, rotate_time=4), dict(state=E.REFINE_INIT, use_apd=True, rotate_time=2)):
    case = T.build_case(97, 71, 3, iters=2, **kw)
    apd = T.make_product(case)
    apd.RunPatchMatch()
    st = T.product_state(apd)
    print("case", kw, float(np.nanmean(st["planes"][..., 3])))
    apd.close()
sc = make_scene(1010, 90, 2)
s = P.Scene(sc["images"].numpy(), sc["cameras"], P.ring_pairs(3, 2))
s.Run()
print("scene", float(s.Depth(0).mean()))
from apd_mvs_b200 import fusion as F
import fusion_tools as FT
fu = F.Fusion(3, 1010, 90)
bgr = FT.colour_images(sc["images"].numpy())
for v in range(3):
    fu.SetView(v, bgr[v], sc["cameras"][v], s.Depth(v), s.Normal(v), s.States(v))
for r, ss in P.ring_pairs(3, 2):
    fu.AddProblem(r, ss)
xyz, col = fu.RunFusion()
print("fusion", len(xyz), fu.Timing()["max_rounds"])
for variant in ("tat_intermediate", "tat_advanced"):
    xyz, col = fu.RunFusion(variant)
    print("fusion", variant, len(xyz))
fu.close()
s.close()
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitize_$tool.log
done
