"""Shared helpers for the GPU parity tests: run the product (libapd_b200.so through the C-ABI) and
the reference oracle (oracle/_ref/libapd_ref.so = the reference's own APD.cu) on identical inputs
and diff the device state after each of the reference's launches."""
from __future__ import annotations

import numpy as np

from apd_mvs_b200 import engine as E
from apd_mvs_b200.scene import make_scene, make_priors, CURAND_SEED


def stage_names(iters: int):
    names = ["K1 rng", "K2 nearest", "K3 anchors", "K4 demote", "K5 init"]
    for i in range(iters):
        names += [f"it{i} K6 strong black", f"it{i} K7 strong red", f"it{i} K8 fit", f"it{i} K9 weak black", f"it{i} K10 weak red"]
    names += ["K11 depth+normal", "K12 filter black", "K13 filter red", "K14 classify", "K15 local refine"]
    return names


def clone_params(p: E.PatchMatchParams) -> E.PatchMatchParams:
    q = E.PatchMatchParams()
    import ctypes as C
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    return q


def build_case(W, H, n_src, state=E.FIRST_INIT, use_apd=False, geom=False, iters=1, device="cpu", seed=None,
               rotate_time=4, ransac_threshold=0.00625, weak_peak_radius=4, top_k=4):
    kw = {} if seed is None else {"seed": seed}
    scene = make_scene(W, H, n_src, device=device, **kw)
    params = E.default_params(max_iterations=iters, state=state, use_APD=1 if use_apd else 0,
                              geom_consistency=1 if geom else 0, rotate_time=rotate_time,
                              ransac_threshold=ransac_threshold, weak_peak_radius=weak_peak_radius, top_k=top_k)
    case = {"scene": scene, "params": params, "images": scene["images"].cpu().numpy(), "cameras": scene["cameras"],
            "depths": None, "planes": None, "views": None, "states": None}
    if state != E.FIRST_INIT or use_apd or geom:
        pri = make_priors(scene)
        if state != E.FIRST_INIT:
            case["planes"], case["views"] = pri["planes"], pri["views"]
        if use_apd:
            case["states"] = pri["states"]
        if geom:
            case["depths"] = pri["depths"]
    return case


def make_product(case, seed=CURAND_SEED) -> E.APD:
    pb = E.Problem(case["images"], case["cameras"], clone_params(case["params"]), depths=case["depths"],
                   planes=case["planes"], views=case["views"], states=case["states"], seed=seed)
    apd = E.APD(pb)
    apd.InuputInitialization()
    apd.CudaSpaceInitialization()
    apd.SetDataPassHelperInCuda()
    return apd


def make_reference(case, seed=CURAND_SEED):
    from oracle.ref_binding import RefAPD
    p = clone_params(case["params"])
    cams = case["cameras"]
    p.depth_min = float(np.float32(cams[0]["depth_min"]) * np.float32(0.6))
    p.depth_max = float(np.float32(cams[0]["depth_max"]) * np.float32(1.2))
    return RefAPD(case["images"], cams, p, depths=case["depths"], planes=case["planes"], views=case["views"],
                  states=case["states"], seed=seed)


def product_state(apd: E.APD):
    return {"planes": apd.GetPlaneHypotheses(), "costs": apd.GetCosts(), "views": apd.GetSelectedViews(),
            "states": apd.GetPixelStates(), "view_weights": apd.GetViewWeights(), "rng": apd.GetRng()}


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def diff_state(mine, ref, fields=("planes", "costs", "views", "states", "view_weights", "rng")):
    """Fraction of pixels whose value differs bit-for-bit, per field (NaNs with equal bits are equal)."""
    out = {}
    for f in fields:
        a, b = _bits(np.ascontiguousarray(mine[f])), _bits(np.ascontiguousarray(ref[f]))
        ne = a != b
        if ne.ndim > 2:
            ne = ne.reshape(ne.shape[0], ne.shape[1], -1).any(axis=-1)
        out[f] = float(ne.mean())
    return out


def depth_stats(mine_planes, ref_planes, depth_min, depth_max):
    dm, dr = mine_planes[..., 3].astype(np.float64), ref_planes[..., 3].astype(np.float64)
    ok = np.isfinite(dm) & np.isfinite(dr)
    rel = np.abs(dm - dr) / np.maximum(np.abs(dr), 1e-12)
    return {"depth_L1_norm": float(np.abs(dm - dr)[ok].mean() / (depth_max - depth_min)),
            "frac_rel_le_1e-4": float((rel[ok] <= 1e-4).mean()), "valid": float(ok.mean())}


def random_config(rng):
    """One random configuration of the sweep used by tests/tools/parity_fuzz.py and tests/test_parity_gpu.py."""
    W, H = int(rng.integers(40, 260)), int(rng.integers(40, 200))
    S = int(rng.choice([1, 2, 3, 4, 5, 7, 9, 12, 17, 31]))
    state = int(rng.choice([E.FIRST_INIT, E.REFINE_INIT, E.REFINE_ITER]))
    use_apd = bool(rng.integers(0, 2)) and state != E.FIRST_INIT
    geom = bool(rng.integers(0, 2)) and state == E.REFINE_ITER
    kw = dict(state=state, use_apd=use_apd, geom=geom, iters=int(rng.integers(1, 4)), rotate_time=int(rng.choice([1, 2, 4])),
              top_k=int(rng.choice([1, 2, 4, 5])), weak_peak_radius=int(rng.choice([2, 4, 6])),
              ransac_threshold=float(rng.choice([0.005, 0.00625, 0.00875])), seed=int(rng.integers(1, 1 << 30)))
    if S > 12:
        W, H = min(W, 120), min(H, 90)
    return W, H, S, kw, int(rng.integers(1, 1 << 40))


def final_diff(case, curand_seed):
    """Bit-difference fractions of the final outputs, product vs reference oracle."""
    ref = make_reference(case, seed=curand_seed); ref.run(); rp, rs, rv = ref.outputs(); ref.close()
    apd = make_product(case, seed=curand_seed); apd.RunPatchMatch(); mine = product_state(apd); apd.close()
    return diff_state({"planes": mine["planes"], "states": mine["states"], "views": mine["views"]},
                      {"planes": rp, "states": rs, "views": rv}, fields=("planes", "states", "views"))
