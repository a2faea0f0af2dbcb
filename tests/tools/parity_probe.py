"""GPU diagnostic (run under gpurun): stage-by-stage bit diff of the product against the reference
oracle, then per-stage timings of both on a larger case. Writes gpurun_out/parity_probe.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_tools as T
from apd_mvs_b200 import engine as E

out = {"cases": []}
os.makedirs("gpurun_out", exist_ok=True)

def run_case(name, W, H, S, iters, stages="all", **kw):
    case = T.build_case(W, H, S, iters=iters, device="cuda", **kw)
    names = T.stage_names(iters)
    ref = T.make_reference(case)
    snaps = list(range(len(names)))
    ref.run(snapshots=snaps)
    apd = T.make_product(case)
    rows = []
    for s in (snaps if stages == "all" else stages):
        apd.RunPatchMatch(stage_end=s)
        d = T.diff_state(T.product_state(apd), ref.get(s))
        rows.append({"stage": s, "name": names[s], **d})
        print(f"[{name}] stage {s:2d} {names[s]:22s} " + " ".join(f"{k}={v:.6f}" for k, v in d.items()), flush=True)
    if case["params"].use_APD:
        apd.RunPatchMatch(stage_end=3)
        anchors, nearest, reliable, fit = apd.GetAnchors()
        comp, nmap, rnear, rrel, rfit, wc = ref.anchors()
        weak = case["states"] == 0
        ra = comp[nmap[weak]]
        ma = anchors[weak]
        print(f"[{name}] weak px {weak.sum()} anchors mismatch frac {np.mean((ra != ma).any(axis=(1, 2))):.6f} nearest mismatch {np.mean((rnear[weak] != nearest[weak]).any(-1)):.6f} reliable mismatch {np.mean(rrel[weak] != reliable[weak]):.6f}", flush=True)
        bad = np.where((ra != ma).any(axis=(1, 2)))[0][:3]
        for b in bad: print("  ref", ra[b].tolist(), "mine", ma[b].tolist())
    apd.RunPatchMatch()
    mine = T.product_state(apd)
    rp, rs, rv = ref.outputs()
    fin = T.diff_state({"planes": mine["planes"], "states": mine["states"], "views": mine["views"]},
                       {"planes": rp, "states": rs, "views": rv}, fields=("planes", "states", "views"))
    ds = T.depth_stats(mine["planes"], rp, apd.GetDepthMin(), apd.GetDepthMax())
    print(f"[{name}] FINAL {fin} {ds}", flush=True)
    out["cases"].append({"name": name, "W": W, "H": H, "S": S, "iters": iters, "stages": rows, "final": fin, "depth": ds,
                         "my_ms": apd.StageMs().tolist(), "ref_ms": ref.stage_ms().tolist()})
    apd.close(); ref.close()

def time_case(name, W, H, S, iters, reps=2, **kw):
    case = T.build_case(W, H, S, iters=iters, device="cuda", **kw)
    names = T.stage_names(iters)
    ref = T.make_reference(case)
    ref.run()
    ref_ms = ref.stage_ms()
    rp, rs, rv = ref.outputs()
    ref.close()
    apd = T.make_product(case)
    for _ in range(reps):
        t0 = time.time(); apd.RunPatchMatch(); t1 = time.time()
    my_ms = apd.StageMs()
    mine = T.product_state(apd)
    fin = T.diff_state({"planes": mine["planes"], "states": mine["states"], "views": mine["views"]},
                       {"planes": rp, "states": rs, "views": rv}, fields=("planes", "states", "views"))
    print(f"[{name}] final diff {fin}; wall {1e3*(t1-t0):.1f} ms")
    for i, n in enumerate(names):
        r = ref_ms[i] if i < len(ref_ms) else float('nan')
        print(f"[{name}] {i:2d} {n:22s} ref {r:10.3f} ms   mine {my_ms[i]:10.3f} ms   x{(r/my_ms[i] if my_ms[i] > 0 else 0):6.2f}", flush=True)
    out["cases"].append({"name": name, "W": W, "H": H, "S": S, "iters": iters, "final": fin,
                         "my_ms": my_ms.tolist(), "ref_ms": ref_ms.tolist()})
    apd.close()

which = sys.argv[1:] or ["small", "mid", "time"]
if "small" in which:
    run_case("cfg1 256x256 S1", 256, 256, 1, 1)
if "mid" in which:
    run_case("mid 320x240 S4 it2", 320, 240, 4, 2)
    run_case("odd 203x131 S3 it1", 203, 131, 3, 1)
if "refine" in which:
    run_case("refine_init 320x240 S4", 320, 240, 4, 1, state=E.REFINE_INIT)
    run_case("refine_iter geom 320x240 S4", 320, 240, 4, 1, state=E.REFINE_ITER, geom=True)
if "apd" in which:
    run_case("apd refine_iter geom 320x240 S4", 320, 240, 4, 2, state=E.REFINE_ITER, geom=True, use_apd=True)
    run_case("apd refine_init 320x240 S4 rot2", 320, 240, 4, 1, state=E.REFINE_INIT, use_apd=True, rotate_time=2, ransac_threshold=0.00875, weak_peak_radius=6)
if "time" in which:
    time_case("time 1024x768 S9 it3", 1024, 768, 9, 3)
if "cfg3" in which:
    time_case("cfg3 6221x4146 S9 apd+geom it3", 6221, 4146, 9, 3, reps=1, state=E.REFINE_ITER, geom=True, use_apd=True)
if "cfg3s" in which:
    time_case("cfg3-small 1555x1036 S9 apd+geom it3", 1555, 1036, 9, 3, reps=2, state=E.REFINE_ITER, geom=True, use_apd=True)
if "cfg2" in which:
    time_case("cfg2 3111x2074 S9 it3", 3111, 2074, 9, 3, reps=2)
json.dump(out, open("gpurun_out/parity_probe.json", "w"), indent=1)
