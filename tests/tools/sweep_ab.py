"""GPU diagnostic: the two depth-sweep kernels (APD_SWEEP_IMPL=old / default) on one small case, in one process each;
prints where their pixel states / depths differ. Usage: python tests/tools/sweep_ab.py W H S"""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 4 and sys.argv[4] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_tools as T
    W, H, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    case = T.build_case(W, H, S, device="cuda", iters=1)
    apd = T.make_product(case); apd.RunPatchMatch()
    np.savez(sys.argv[5], planes=apd.GetPlaneHypotheses(), states=apd.GetPixelStates(), views=apd.GetSelectedViews(), vw=apd.GetViewWeights())
    apd.close(); sys.exit(0)
W, H, S = sys.argv[1:4]
outs = []
for impl in ("old", "new"):
    env = dict(os.environ); 
    if impl == "old": env["APD_SWEEP_IMPL"] = "old"
    f = f"/tmp/sweep_ab_{impl}.npz"
    subprocess.run([sys.executable, __file__, W, H, S, "child", f], env=env, check=True)
    outs.append(np.load(f))
a, b = outs
W, H = int(W), int(H)
ds = (a["states"].reshape(H, W) != b["states"].reshape(H, W))
dp = (a["planes"].reshape(H, W, 4).view(np.uint32) != b["planes"].reshape(H, W, 4).view(np.uint32)).any(-1)
views = a["views"].reshape(H, W)
nv = np.array([bin(int(v)).count("1") for v in views.ravel()]).reshape(H, W)
print("states differ", int(ds.sum()), "planes differ", int(dp.sum()), "of", W * H)
ys, xs = np.nonzero(ds)
print("state diffs: x range", xs.min() if len(xs) else None, xs.max() if len(xs) else None, "y range", ys.min() if len(ys) else None, ys.max() if len(ys) else None)
print("first 12 state diffs (x, y, old, new, nsel):", [(int(x), int(y), int(a["states"].reshape(H, W)[y, x]), int(b["states"].reshape(H, W)[y, x]), int(nv[y, x])) for y, x in list(zip(ys, xs))[:12]])
print("x mod 8 histogram of state diffs:", np.bincount(xs % 8, minlength=8).tolist(), " y mod 4:", np.bincount(ys % 4, minlength=4).tolist())
ys, xs = np.nonzero(dp)
print("first 12 plane diffs (x, y, old w, new w):", [(int(x), int(y), float(a["planes"].reshape(H, W, 4)[y, x, 3]), float(b["planes"].reshape(H, W, 4)[y, x, 3])) for y, x in list(zip(ys, xs))[:12]])
print("nsel histogram all:", np.bincount(nv.ravel(), minlength=int(S) + 1).tolist(), "at state diffs:", np.bincount(nv[ds], minlength=int(S) + 1).tolist())
