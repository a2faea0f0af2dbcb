"""Diagnostic: K3 (anchors, RNG position, reliability) after stage 2 from the default build and from APD_K3_GENERIC=1,
on the full-size cfg3 case; prints where they differ.   python tools/k3_diff.py dump <npz> | python tools/k3_diff.py cmp a.npz b.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if sys.argv[1] == "dump":
    import parity_tools as T
    from apd_mvs_b200 import engine as E
    W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (6221, 4146)
    case = T.build_case(W, H, 9, iters=3, device="cuda", state=E.REFINE_ITER, geom=True, use_apd=True)
    apd = T.make_product(case)
    apd.RunPatchMatch(stage_end=2)
    anchors, nearest, reliable, _ = apd.GetAnchors()
    np.savez(sys.argv[2], anchors=anchors, reliable=reliable, rng=apd.GetRng(), states=case["states"])
    apd.close()
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    weak = a["states"] == 0
    da = (a["anchors"] != b["anchors"]).any(axis=(2, 3)) & weak
    dr = (a["rng"] != b["rng"]).any(axis=2) & weak
    dl = (a["reliable"] != b["reliable"]) & weak
    print("weak px", int(weak.sum()), "anchor diffs", int(da.sum()), "rng diffs", int(dr.sum()), "reliable diffs", int(dl.sum()))
    ys, xs = np.nonzero(da | dr | dl)
    print("first differing pixels (x, y):", list(zip(xs[:12].tolist(), ys[:12].tolist())))
    if len(ys):
        print("bbox x", xs.min(), xs.max(), "y", ys.min(), ys.max())
        for x, y in list(zip(xs[:4].tolist(), ys[:4].tolist())):
            print((x, y), "A", a["anchors"][y, x].tolist(), a["rng"][y, x].tolist(), int(a["reliable"][y, x]))
            print((x, y), "B", b["anchors"][y, x].tolist(), b["rng"][y, x].tolist(), int(b["reliable"][y, x]))
