"""Generates the golden fixtures from the REFERENCE ITSELF (oracle/_ref/libapd_ref.so = the reference's
own APD.cu recompiled for sm_100, see oracle/ref_wrapper.cu). Needs a GPU:

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Each fixture stores the INPUTS (images, cameras, priors) next to the reference's device state after
selected launches, so tests never depend on re-generating inputs bit-identically on another machine.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_tools as T
from apd_mvs_b200 import engine as E

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)


def params_bytes(p):
    import ctypes as C
    return np.frombuffer(C.string_at(C.byref(p), C.sizeof(p)), dtype=np.uint8).copy()


def make(name, W, H, S, iters, stages, **kw):
    case = T.build_case(W, H, S, iters=iters, device="cpu", **kw)
    ref = T.make_reference(case)
    ref.run(snapshots=stages)
    d = {"images": case["images"], "cameras": case["cameras"].view(np.uint8).reshape(S + 1, 112),
         "params": params_bytes(case["params"]), "seed": np.uint64(T.CURAND_SEED), "stages": np.array(stages, np.int32)}
    for k in ("depths", "planes", "views", "states"):
        if case[k] is not None:
            d["in_" + k] = case[k]
    for s in stages:
        st = ref.get(s)
        d[f"s{s}_planes"], d[f"s{s}_costs"], d[f"s{s}_views"], d[f"s{s}_states"] = st["planes"], st["costs"], st["views"], st["states"]
        d[f"s{s}_vw"] = st["view_weights"][..., :8].copy()
        if s == 0 or s == stages[-1]:
            d[f"s{s}_rng"] = st["rng"]
    d["planes"], d["states"], d["views"] = ref.outputs()
    if case["params"].use_APD:
        comp, nmap, nearest, reliable, fit, wc = ref.anchors()
        d["anchors_compact"], d["anchors_map"], d["nearest"], d["reliable"], d["fit_planes"] = comp[:wc], nmap, nearest, reliable, fit
    ref.close()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, {k: v.shape for k, v in d.items() if hasattr(v, "shape") and v.ndim > 1})


make("strong_first_64x48_s2", 64, 48, 2, 1, [0, 4, 5, 6, 10, 12, 13, 14])
make("strong_geom_64x48_s3", 64, 48, 3, 1, [4, 6, 14], state=E.REFINE_ITER, geom=True)
make("strong_refineinit_48x40_s2", 48, 40, 2, 2, [4, 11, 19], state=E.REFINE_INIT)
make("smoke_128x96", 128, 96, 2, 1, [14])
if "apd" in sys.argv:
    make("apd_geom_96x72_s3", 96, 72, 3, 1, [1, 2, 3, 4, 6, 7, 9, 14], state=E.REFINE_ITER, geom=True, use_apd=True)
    make("apd_init_96x72_s3_rot2", 96, 72, 3, 1, [2, 3, 9, 14], state=E.REFINE_INIT, use_apd=True, rotate_time=2, ransac_threshold=0.00875, weak_peak_radius=6)
