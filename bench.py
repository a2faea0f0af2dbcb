#!/usr/bin/env python
"""bench.py — headline benchmark of the PatchMatch hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1|mid|cfg5|cfg3s|cfg4]

metric   Mpixels*views/s per PatchMatch iteration = W*H*S / t_iter / 1e6, t_iter = GPU time of one
         iteration of the loop APD.cu:2443-2457 (strong black + strong red + fit plane + weak black + weak
         red), averaged over the iterations of all timed steps.
workload default cfg3 = BASELINE.json configs[2]: 6221x4146, 1 ref x 9 src, REFINE_ITER with adaptive patch
         deformation ON and the geometric-consistency term, the largest single-GPU configuration and the shape
         north_star's target names. cfg2 (configs[1], all-STRONG FIRST_INIT) is printed as `secondary.cfg2`.
step     one full RunPatchMatch (all launches of APD.cu:2409-2471) on one reference view, inputs resident in HBM.
e2e      the same metric through the C-ABI with HOST buffers: every step uploads images, cameras, depth maps and
         priors from pinned host memory, runs, and reads planes/states/views back (what ProcessProblem,
         main.cpp:95-124, does per view); `incl_create_destroy` adds apd_create/apd_destroy per step (the lifecycle
         the reference arm pays).
roofline contract definition: algorithmic bytes = 8 B per NCC tap of the REFERENCE algorithm, counted
         from the input masks (WEAK/STRONG counts per colour, anchor counts, selected-view popcounts), for k_strong,
         k_weak and k_sweep; `tex` = the physical figure: executed thread-level texture fetches (ncu, profiles/) /
         time / (4 fetches/clk/SM x SMs x SM clock).
N > 1    reference views sharded one per rank (weak scaling, no data-path collective): rank 0 renders the view
         ring and broadcasts images, cameras and depth maps once over NCCL (apd_mvs_b200/shard.py); value = all
         ranks' units / max time. `secondary.cfg4` = the sharded pass schedule of BASELINE configs[3] (strong scaling).
--impl reference   times the reference's own CUDA build (oracle/_ref/libapd_ref.so = unmodified APD.cu
         recompiled for sm_100; the reference has no CPU path) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FIRST_INIT, REFINE_INIT, REFINE_ITER = 0, 1, 2
WORKLOADS = {
    "cfg3": dict(W=6221, H=4146, S=9, iters=3, state=REFINE_ITER, use_apd=True, geom=True, rotate_time=4, ransac_threshold=0.00625, weak_peak_radius=4,
                 desc="synthetic ETH3D-full-res 6221x4146, 1 ref x 9 src, REFINE_ITER, deformation ON + geom-consistency, 3 iters (BASELINE.json configs[2])"),
    "cfg2": dict(W=3111, H=2074, S=9, iters=3, state=FIRST_INIT,
                 desc="synthetic ETH3D-half-res shape 3111x2074, 1 ref x 9 src views, all STRONG, 3 iters (BASELINE.json configs[1])"),
    "cfg1": dict(W=256, H=256, S=1, iters=1, state=FIRST_INIT, desc="synthetic 2-view 256x256, 1 iter (BASELINE.json configs[0])"),
    "mid": dict(W=1024, H=768, S=9, iters=3, state=FIRST_INIT, desc="synthetic 1024x768, 1 ref x 9 src, 3 iters (development size)"),
    "cfg5": dict(W=4096, H=4096, S=16, iters=8, state=FIRST_INIT,
                 desc="synthetic 4096x4096, 1 ref x 16 src views, all STRONG, 8 iters (BASELINE.json configs[4])"),
    "cfg3s": dict(W=1555, H=1036, S=9, iters=3, state=REFINE_ITER, use_apd=True, geom=True, rotate_time=4, ransac_threshold=0.00625, weak_peak_radius=4,
                  desc="cfg3 at quarter resolution 1555x1036 (development size)"),
}
ALG_BYTES_PER_TAP = 8                       # one fp32 reference sample + one fp32 source sample
TAPS_STRONG = 14 * 36                       # 14 hypotheses x 36 taps per strong pixel*view*iter


class ClockSampler:
    def __init__(self, device_index: int):
        self.proc = None
        self.lines = []
        self.idx = device_index
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def rd():
            for ln in self.proc.stdout:
                self.lines.append(ln.strip())
        self.thread = threading.Thread(target=rd, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the samples taken under load (upper half of the observed clocks)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def crc_outputs(np, planes, states, views):
    return {"planes": zlib.crc32(np.ascontiguousarray(planes).tobytes()), "states": zlib.crc32(np.ascontiguousarray(states).tobytes()),
            "views": zlib.crc32(np.ascontiguousarray(views).tobytes())}


def build_inputs(wl, rank, world, local, impl, torch, np, dist):
    """Renders the view ring on rank 0, broadcasts it (the single setup collective) and returns this rank's
    problem: its reference view and the S next views of the ring, plus priors when the workload refines."""
    from apd_mvs_b200 import shard
    from apd_mvs_b200.scene import make_scene, make_priors, CAMERA_DTYPE
    W, H, S = wl["W"], wl["H"], wl["S"]
    need_priors = wl["state"] != FIRST_INIT or wl.get("use_apd") or wl.get("geom")
    n_units = world if impl == "ours" else 1
    n_views = S + n_units
    dev = f"cuda:{local}"
    scene0 = None
    if rank == 0:
        scene0 = make_scene(W, H, n_views - 1, device=dev)
        images_all, depth_all = scene0["images"], scene0["depth"]
        cams_all = torch.from_numpy(scene0["cameras"].view(np.uint8).reshape(n_views, 112).copy()).to(dev)
    else:
        images_all = torch.empty((n_views, H, W), dtype=torch.float32, device=dev)
        depth_all = torch.empty((n_views, H, W), dtype=torch.float32, device=dev) if need_priors else None
        cams_all = torch.empty((n_views, 112), dtype=torch.uint8, device=dev)
    bcast_ms = 0.0
    if world > 1 and impl == "ours":
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        shard.broadcast_inputs(images_all, cams_all, 0)
        if need_priors:
            dist.broadcast(depth_all, 0)      # what the previous pass of the other ranks produced (all-gather in the schedule)
        torch.cuda.synchronize()
        bcast_ms = 1e3 * (time.perf_counter() - t0)
    order = shard.view_order(rank, S, n_views)
    cams = cams_all[order].cpu().numpy().copy().view(CAMERA_DTYPE).reshape(-1)
    case = {"images": images_all[order].contiguous(), "cameras": cams, "depths": None, "planes": None, "views": None, "states": None}
    if need_priors:
        if rank == 0:
            sc = scene0
        else:   # this rank's reference view: normal + textureless mask (images and depth maps came from rank 0)
            sc = make_scene(W, H, n_views - 1, device=dev, only=[order[0]], ref_index=order[0])
        sc = dict(sc); sc["depth"] = depth_all
        pri = make_priors(sc, order=order)
        if wl["state"] != FIRST_INIT:
            case["planes"], case["views"] = pri["planes"], pri["views"]
        if wl.get("use_apd"):
            case["states"] = pri["states"]
        if wl.get("geom"):
            case["depths"] = pri["depths"]
    del images_all, depth_all, scene0
    torch.cuda.empty_cache()
    return case, bcast_ms


def make_params(E, wl):
    return E.default_params(max_iterations=wl["iters"], state=wl["state"], use_APD=1 if wl.get("use_apd") else 0,
                            geom_consistency=1 if wl.get("geom") else 0, rotate_time=wl.get("rotate_time", 4),
                            ransac_threshold=wl.get("ransac_threshold", 0.005), weak_peak_radius=wl.get("weak_peak_radius", 2))


def contract_bytes(np, wl, states_k4, anchors, views_final):
    """Algorithmic bytes per launch of the three NCC kernels, from the input masks. states_k4: pixel states
    after K4 (what the propagation kernels see); anchors [H,W,9,2]; views_final: selected-view bitmasks K14/K15 read."""
    H, W = states_k4.shape
    S = wl["S"]
    half_rows = 32 * ((H // 2 + 15) // 16)               # rows the reference's half launch reaches (APD.cu:2400-2403)
    yy, xx = np.mgrid[0:H, 0:W]
    reach = yy < half_rows
    black = ((xx + yy) & 1) == 0
    weak = states_k4 == 0
    out = {}
    n_strong = [int((reach & ~weak & (black if c == 0 else ~black)).sum()) for c in (0, 1)]
    out["k_strong"] = {"pixels_per_launch": sum(n_strong) / 2.0, "bytes_per_launch": sum(n_strong) / 2.0 * S * TAPS_STRONG * ALG_BYTES_PER_TAP}
    if anchors is not None and weak.any():
        a = (anchors[..., 1:, 0] != -1).sum(-1)                                  # valid anchors per pixel (slots 1..8)
        taps = (15 * (36 + 9 * a) + 36) * (weak & reach)                         # per WEAK pixel*view*iter
        n_weak = int((weak & reach).sum())
        out["k_weak"] = {"pixels_per_launch": n_weak / 2.0, "mean_anchors": float(a[weak & reach].mean()) if n_weak else 0.0,
                         "bytes_per_launch": float(taps.sum()) / 2.0 * S * ALG_BYTES_PER_TAP}
    pop = np.zeros((H, W), np.int64)
    for v in range(S):
        pop += (views_final >> np.uint32(v)) & 1
    interior = np.zeros((H, W), bool); interior[6:-6, 6:-6] = True
    out["k_sweep"] = {"pixel_views_k14": int(pop[interior].sum()), "pixel_views_k15": int(pop.sum()),
                      "bytes_per_launch": float(pop[interior].sum()) * 62 * 36 * ALG_BYTES_PER_TAP + float(pop.sum()) * 12 * 36 * ALG_BYTES_PER_TAP}
    return out


def run_ours(wl_name, args, rank, world, local, torch, np, dist, full=True):
    import ctypes as C
    import parity_tools as T
    from apd_mvs_b200 import engine as E
    wl = WORKLOADS[wl_name]
    W, H, S, iters = wl["W"], wl["H"], wl["S"], wl["iters"]
    npx = W * H
    case, bcast_ms = build_inputs(wl, rank, world, local, "ours", torch, np, dist)
    case["params"] = make_params(E, wl)
    seed = 1234567 + rank
    names = T.stage_names(iters)
    iter_idx = [i for i, n in enumerate(names) if n.startswith("it")]
    idx = {"k_strong": [i for i, n in enumerate(names) if "strong" in n], "k_weak": [i for i, n in enumerate(names) if "weak" in n],
           "k_sweep": [i for i, n in enumerate(names) if "K14" in n]}
    sampler = ClockSampler(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    pb = E.Problem(case["images"], case["cameras"], T.clone_params(case["params"]), depths=case["depths"], planes=case["planes"],
                   views=case["views"], states=case["states"], seed=seed, device=local)
    apd = E.APD(pb)
    apd.InuputInitialization(); apd.CudaSpaceInitialization(); apd.SetDataPassHelperInCuda()
    steps, warmup = (args.steps, args.warmup) if full else (3, 2)
    for _ in range(warmup):
        apd.RunPatchMatch()
    barrier()
    sampler.start()
    stage_ms = np.zeros(len(names))
    t0 = time.perf_counter()
    for _ in range(steps):
        apd.RunPatchMatch()                      # device-resident inputs; events on the engine's own stream
        stage_ms += apd.StageMs()
    barrier()
    total_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    res = {"launches": apd.LaunchCount() * steps, "dev_step_ms": float(stage_ms.sum() / steps),
           "iter_ms": float(stage_ms[iter_idx].sum() / (steps * iters)), "total_ms": total_ms, "steps": steps,
           "stage_ms": (stage_ms / steps).round(3).tolist(), "clocks": clocks, "bcast_ms": bcast_ms,
           "kernel_ms": {k: float(stage_ms[v].sum() / (steps * len(v))) if v else 0.0 for k, v in idx.items()}}
    planes, states, views = apd.GetPlaneHypotheses(), apd.GetPixelStates(), apd.GetSelectedViews()
    res["crc"] = crc_outputs(np, planes, states, views)
    if not full:
        apd.close()
        return res, case, None
    # ---- e2e through the C-ABI with pinned HOST buffers
    L = E.lib()
    hp = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    images_host = case["images"].cpu().pin_memory()
    img_ptrs = (C.c_void_p * (S + 1))(*[images_host.data_ptr() + i * npx * 4 for i in range(S + 1)])
    cam_arr = np.ascontiguousarray(case["cameras"])
    h2d = (S + 1) * npx * 4 + 112 * (S + 1)
    dep_host = dep_ptrs = pl_host = vw_host = st_host = None
    if case["depths"] is not None:
        dep_host = hp(case["depths"])
        dep_ptrs = (C.c_void_p * (S + 1))(*[dep_host.data_ptr() + i * npx * 4 for i in range(S + 1)])
        h2d += (S + 1) * npx * 4
    if case["planes"] is not None:
        pl_host, vw_host = hp(case["planes"]), hp(case["views"].view(np.int32))
        h2d += npx * 20
    if case["states"] is not None:
        st_host = hp(case["states"]); h2d += npx
    out_planes = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    out_states = torch.empty((H, W), dtype=torch.uint8).pin_memory()
    out_views = torch.empty((H, W), dtype=torch.int32).pin_memory()
    vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)

    def e2e_step(h):
        # asynchronous upload mode of the C-ABI (apd_set_upload_mode): the calls enqueue their copies from the pinned host
        # buffers and apd_run makes every launch wait for the inputs it reads - cameras + priors first (the first launches
        # read them), then the image and depth stacks, whose transfer overlaps K1..K4
        rc = L.apd_set_upload_mode(h, 1)
        rc |= L.apd_set_cameras(h, C.c_void_p(cam_arr.ctypes.data))
        if pl_host is not None or st_host is not None:
            rc |= L.apd_set_priors(h, vp(pl_host), vp(vw_host), vp(st_host))
        rc |= L.apd_set_images(h, img_ptrs, W * 4)
        if dep_host is not None:
            rc |= L.apd_set_depths(h, dep_ptrs, W * 4)
        rc |= L.apd_run(h)
        rc |= L.apd_get_planes(h, vp(out_planes)); rc |= L.apd_get_states(h, vp(out_states)); rc |= L.apd_get_views(h, vp(out_views))
        if rc:
            raise RuntimeError("C-ABI call failed in the e2e step: " + (L.apd_last_error(h) or b"").decode())
    e2e_step(apd._h)
    barrier()
    n_e2e = max(2, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step(apd._h)
    barrier()
    res["e2e_ms"] = 1e3 * (time.perf_counter() - t0) / n_e2e
    res["e2e_crc_equal"] = crc_outputs(np, out_planes.numpy(), out_states.numpy(), out_views.numpy().view(np.uint32)) == res["crc"]
    res["h2d"], res["d2h"] = h2d, npx * 21
    # ---- masks for the contract roofline: states after K4, anchors, final selected views
    masks = None
    if rank == 0:
        anchors = None
        if wl.get("use_apd"):
            apd.RunPatchMatch(stage_end=3)
            states_k4 = apd.GetPixelStates()
            anchors = apd.GetAnchors()[0]
        else:
            states_k4 = np.ones((H, W), np.uint8)
        masks = contract_bytes(np, wl, states_k4, anchors, views)
        del anchors
    apd.close()
    # ---- the lifecycle the reference pays per view (ProcessProblem): create + upload + run + read back + destroy
    def lifecycle():
        h = C.c_void_p(None)
        p = T.clone_params(pb.params)
        if L.apd_create(C.byref(h), local, W, H, S + 1, C.byref(p), seed):
            raise RuntimeError("apd_create failed")
        e2e_step(h)
        L.apd_destroy(h)
    lifecycle()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        lifecycle()
    barrier()
    res["e2e_lifecycle_ms"] = 1e3 * (time.perf_counter() - t0) / 2
    return res, case, masks


def run_reference(wl_name, args, local, torch, np, full=True):
    import parity_tools as T
    from apd_mvs_b200 import engine as E
    wl = WORKLOADS[wl_name]
    iters = wl["iters"]
    case, _ = build_inputs(wl, 0, 1, local, "reference", torch, np, None)
    case["images"] = case["images"].cpu().numpy()
    case["params"] = make_params(E, wl)
    names = T.stage_names(iters)
    iter_idx = [i for i, n in enumerate(names) if n.startswith("it")]
    idx = {"k_strong": [i for i, n in enumerate(names) if "strong" in n], "k_weak": [i for i, n in enumerate(names) if "weak" in n],
           "k_sweep": [i for i, n in enumerate(names) if "K14" in n or "K15" in n]}
    sampler = ClockSampler(local)
    steps, warmup = (args.steps, args.warmup) if full else (2, 1)
    stage_acc, e2e_list, crc = None, [], None
    for it in range(warmup + steps):
        if it == warmup:
            torch.cuda.synchronize(); sampler.start(); t_all = time.perf_counter()
        s0 = time.perf_counter()
        ref = T.make_reference(case, seed=1234567)       # construct + upload  (APD.cpp:356-699)
        ref.run()                                        # APD::RunPatchMatch (incl. its own D2H, APD.cu:2490-2492)
        rp, rs, rv = ref.outputs()
        ms = ref.stage_ms()
        ref.close()                                      # ~APD
        if it >= warmup:
            e2e_list.append(1e3 * (time.perf_counter() - s0))
            stage_acc = ms if stage_acc is None else stage_acc + ms
        crc = crc_outputs(np, rp, rs, rv)
    total_ms = 1e3 * (time.perf_counter() - t_all)
    clocks = sampler.stop()
    stage_ms = stage_acc / steps
    res = {"launches": 25 * steps, "dev_step_ms": float(stage_ms.sum()), "iter_ms": float(stage_ms[iter_idx].sum() / iters), "total_ms": total_ms,
           "steps": steps, "stage_ms": np.round(stage_ms[:len(names)], 3).tolist(), "clocks": clocks, "bcast_ms": 0.0,
           "kernel_ms": {k: float(stage_ms[v].sum() / (len(v) if k != "k_sweep" else 1)) if v else 0.0 for k, v in idx.items()},
           "e2e_ms": float(np.mean(e2e_list)), "h2d": 0, "d2h": 0, "crc": crc}
    return res, case, None


def run_cfg4(rank, world, local, torch, np, dist):
    """BASELINE configs[3]: 1920x1080, 32 reference views x 10 source views, the FIRST_INIT pass and one REFINE_ITER +
    geometric-consistency pass through the sharded pass scheduler (apd_mvs_b200.pipeline.ShardedScene over the scene layer
    of include/apd_scene.h): problems round-robin over the ranks, one all-gather of the depth maps between the passes.
    Fixed total work = strong scaling; the line carries the per-rank process / wait / exchange split."""
    from apd_mvs_b200 import pipeline as P, shard
    from apd_mvs_b200.scene import make_scene, CAMERA_DTYPE
    W, H, V, S, iters = 1920, 1080, 32, 10, 3
    dev = f"cuda:{local}"
    if rank == 0:
        sc = make_scene(W, H, V - 1, device=dev)
        images = sc["images"]
        cams_t = torch.from_numpy(sc["cameras"].view(np.uint8).reshape(V, 112).copy()).to(dev)
        del sc
    else:
        images = torch.empty((V, H, W), dtype=torch.float32, device=dev)
        cams_t = torch.empty((V, 112), dtype=torch.uint8, device=dev)
    if world > 1:
        shard.broadcast_inputs(images, cams_t, 0)
    cams = cams_t.cpu().numpy().copy().view(CAMERA_DTYPE).reshape(-1)
    pairs = P.ring_pairs(V, S)
    scene = P.Scene(images, cams, pairs, seed=4242, device=local, round_limit=max(W, H))     # one round at full size
    assert scene.ComputeRoundNum() == 1
    sh = P.ShardedScene(scene, pairs, rank, world, 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sh.run(passes=(0, 1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = torch.tensor([1e3 * (time.perf_counter() - t0)], dtype=torch.float64, device=dev)
    per_rank = torch.zeros((world, 3), dtype=torch.float64, device=dev)
    per_rank[rank, 0], per_rank[rank, 1], per_rank[rank, 2] = sh.process_ms, sh.wait_ms, sh.exchange_ms
    if world > 1:
        shard.max_over_ranks(wall)
        dist.all_reduce(per_rank)
    valid = float((torch.from_numpy(scene.Depth(pairs[sh.my_problems()[0]][0])) > 0).float().mean())
    scene.close()
    del images
    torch.cuda.empty_cache()
    wall_ms = float(wall[0])
    runs = 2 * V
    return {"workload": "cfg4: synthetic Tanks&Temples shape 1920x1080, 32 ref views x 10 src (ring), FIRST_INIT pass + one REFINE_ITER+geom pass, ref views sharded over the ranks (BASELINE.json configs[3])",
            "scaling": "strong", "n_gpus": world, "runs": runs, "wall_ms": round(wall_ms, 1), "runs_per_s": round(runs / (wall_ms * 1e-3), 2),
            "value": round(runs * iters * W * H * S / (wall_ms * 1e-3) / 1e6, 1), "unit": "Mpixels*views/s per PatchMatch iteration, whole schedule wall time (all launches, exchange and waiting included)",
            "per_rank_process_ms": [round(float(x), 1) for x in per_rank[:, 0]], "per_rank_wait_ms": [round(float(x), 1) for x in per_rank[:, 1]],
            "per_rank_exchange_ms": [round(float(x), 1) for x in per_rank[:, 2]], "valid_depth_fraction_first_owned_view": round(valid, 4),
            "collective": "1 broadcast of images+cameras at setup, 1 all_gather of the owned depth maps per pass (NCCL)"}


def parity_check(case, crc, np):
    """One untimed run of the reference oracle on this rank's inputs, compared bit for bit (checker only)."""
    try:
        import parity_tools as T
        from oracle import ref_binding
        if not ref_binding.available():
            return None
        c = dict(case)
        if not isinstance(c["images"], np.ndarray):
            c["images"] = c["images"].cpu().numpy()
        ref = T.make_reference(c, seed=1234567)
        ref.run()
        rp, rs, rv = ref.outputs()
        ref.close()
        return crc_outputs(np, rp, rs, rv) == crc
    except Exception as e:  # pragma: no cover
        return f"unavailable: {e}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    wl = WORKLOADS[args.workload]
    W, H, S, iters = wl["W"], wl["H"], wl["S"], wl["iters"]
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; the product has no CPU fallback"}))
        sys.exit(2)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")

    if args.impl == "reference" and world > 1:
        dist.barrier()                 # brings NCCL up (and its banner out) before anything is printed
        if rank != 0:                  # the reference is a single-GPU program: rank 0 alone runs it
            dist.destroy_process_group()
            return

    import parity_tools as T
    from apd_mvs_b200 import engine as E
    from apd_mvs_b200 import shard

    dev = f"cuda:{local}"
    npx = W * H
    if args.impl == "ours":
        res, case, masks = run_ours(args.workload, args, rank, world, local, torch, np, dist)
    else:
        from oracle import ref_binding
        if not ref_binding.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libapd_ref.so not built (needs /root/reference at build time)"}))
            return
        res, case, masks = run_reference(args.workload, args, local, torch, np)

    # ---- max over ranks
    vals = torch.tensor([res["iter_ms"], res["dev_step_ms"], res["e2e_ms"], res["total_ms"], res.get("e2e_lifecycle_ms", 0.0)]
                        + [res["kernel_ms"][k] for k in ("k_strong", "k_weak", "k_sweep")], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1 and args.impl == "ours":
        # what every rank measured on ITS reference view: the units differ (each view has its own share of WEAK pixels), so
        # max-over-ranks timing includes that imbalance; there is no communication inside a step
        n_weak = float((case["states"] == 0).sum()) if case.get("states") is not None else 0.0
        mine = torch.tensor([res["iter_ms"], res["dev_step_ms"], res["kernel_ms"]["k_weak"], n_weak], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"iter_ms": [round(float(t[0]), 3) for t in allr], "ms_per_step": [round(float(t[1]), 3) for t in allr],
                    "k_weak_ms": [round(float(t[2]), 3) for t in allr], "weak_pixels": [int(t[3]) for t in allr]}
        shard.max_over_ranks(vals)
    iter_ms, dev_step_ms, e2e_ms, total_ms, life_ms, ms_strong, ms_weak, ms_sweep = [float(v) for v in vals.tolist()]
    n_units = world if args.impl == "ours" else 1
    value = n_units * npx * S / (iter_ms * 1e-3) / 1e6
    e2e_value = n_units * npx * S / (e2e_ms / iters * 1e-3) / 1e6

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        prof = {}
        try:   # per-launch ncu counters of the three kernels (dram bytes, executed thread-level texture fetches), committed under profiles/
            prof = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json"))).get(args.workload, {})
        except Exception:
            pass
        kms = {"k_strong": ms_strong, "k_weak": ms_weak, "k_sweep": ms_sweep}
        nlaunch = {"k_strong": 2 * iters, "k_weak": 2 * iters if wl.get("use_apd") else 0, "k_sweep": 1}
        ref_names = {"k_strong": "Black/RedPixelUpdateStrong (K6/K7)", "k_weak": "Black/RedPixelUpdateWeak (K9/K10)", "k_sweep": "DepthToWeak + LocalRefine (K14+K15)"}
        sm_clock = (res["clocks"].get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        kernels = []
        if masks is None and args.impl == "reference":
            try:   # the reference arm reports the same contract bytes: masks cached by the last run of our arm on this workload
                masks = json.load(open(os.path.join(ROOT, "profiles", "contract_masks.json"))).get(args.workload)
            except Exception:
                masks = None
        for k in ("k_strong", "k_weak", "k_sweep"):
            if not nlaunch[k] or kms[k] <= 0 or not masks or k not in masks:
                continue
            b = masks[k]["bytes_per_launch"]
            ach = b / (kms[k] * 1e-3) / 1e9
            ent = {"kernel": k if args.impl == "ours" else ref_names[k], "replaces": ref_names[k], "ms_per_launch": round(kms[k], 3), "launches_per_step": nlaunch[k],
                   "share_of_step": round(kms[k] * nlaunch[k] / dev_step_ms, 4), "algorithmic_bytes_per_launch": b,
                   "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4),
                   "traffic": prof.get(k, {}).get("dram_bytes_per_launch") if args.impl == "ours" else None,
                   "mask": {kk: vv for kk, vv in masks[k].items() if kk != "bytes_per_launch"}}
            fetches = prof.get(k, {}).get("tex_thread_fetches_per_launch") if args.impl == "ours" else None
            if fetches:
                tex_peak = 4.0 * 148 * sm_clock
                ent["tex"] = {"executed_fetches_per_launch": fetches, "fetches_per_s": round(fetches / (kms[k] * 1e-3), 1),
                              "peak_fetches_per_s": tex_peak, "frac": round(fetches / (kms[k] * 1e-3) / tex_peak, 4),
                              "taps_executed_over_contract": round(fetches / (b / ALG_BYTES_PER_TAP), 4),
                              "source": prof.get(k, {}).get("source", "profiles/")}
            kernels.append(ent)
        dom = max(kernels, key=lambda e: e["share_of_step"]) if kernels else None
        roofline = {"bound": "hbm", "kernel": None, "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None}
        if dom:
            roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["frac"],
                        "traffic": dom["traffic"], "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"]}
            if "tex" in dom:
                roofline["tex"] = dom["tex"]
        roofline["peak_source"] = peak_src
        roofline["kernels"] = kernels
        roofline["note"] = ("contract roofline: algorithmic bytes = 8 B per NCC tap of the REFERENCE algorithm (SURVEY §8d), counted from the input masks, "
                            "independent of what the kernel fetches; none of those bytes come from HBM (traffic = ncu dram bytes per launch), exact skips push "
                            "frac above what is fetched. The physical ceiling is the texture unit: `tex.frac` = executed fetches / (4 per clk per SM).")
        if args.impl == "ours" and masks and args.workload in ("cfg3", "cfg2"):
            try:   # cached for the reference arm's line (same workload, same masks)
                path = os.path.join(ROOT, "profiles", "contract_masks.json")
                allm = json.load(open(path)) if os.path.exists(path) else {}
                allm[args.workload] = masks
                json.dump(allm, open(path, "w"), indent=1)
                if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
                    json.dump(allm, open(os.path.join(ROOT, "gpurun_out", "contract_masks.json"), "w"), indent=1)
            except Exception:
                pass
        line = {
            "metric": "Mpixels*views/s per PatchMatch iteration", "value": round(value, 2), "unit": "Mpixels*views/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_step_ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": args.impl,
            "config": {"workload": f"{args.workload}: {wl['desc']}", "width": W, "height": H, "src_views": S, "iters": iters,
                       "ref_views_per_step": n_units, "state": ["FIRST_INIT", "REFINE_INIT", "REFINE_ITER"][wl["state"]],
                       "use_APD": bool(wl.get("use_apd")), "geom_consistency": bool(wl.get("geom")),
                       "l2": "inputs+state larger than L2 (126 MB)" if npx * (4 * (S + 1) + 130) > 126e6 else "fits L2; state rewritten every step",
                       "parallelism": f"ref-views-dp{world}"},
            "iter_ms": round(iter_ms, 3), "kernel_ms": {k: round(v, 3) for k, v in kms.items()}, "wall_ms_per_step": round(total_ms / res["steps"], 3),
            "stage_ms": res["stage_ms"],
            "e2e": {"value": round(e2e_value, 2), "unit": "Mpixels*views/s", "ms_per_call": round(e2e_ms, 3),
                    "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"]},
            "gpu_launches": res["launches"],
            "clocks": res["clocks"],
            "roofline": roofline,
            "output_crc32": res["crc"],
        }
        if args.impl == "ours":
            line["e2e"]["incl_create_destroy"] = {"value": round(n_units * npx * S / (life_ms / iters * 1e-3) / 1e6, 2), "ms_per_call": round(life_ms, 3),
                                                  "note": "apd_create + uploads + run + read-back + apd_destroy per step: the per-view lifecycle of ProcessProblem that the reference arm's e2e pays"}
            line["e2e"]["outputs_equal_device_path"] = res.get("e2e_crc_equal")
        if world > 1:
            line["setup_broadcast_ms"] = round(res["bcast_ms"], 2)
            if per_rank:
                line["per_rank"] = per_rank
                line["per_rank"]["note"] = ("one reference view per rank, no collective inside a step; `value` uses the slowest rank, whose view simply holds more "
                                            "WEAK pixels - the spread below is input imbalance, not communication")
        if args.impl == "reference":
            line["cpu_baseline"] = {"value": round(value, 2), "unit": "Mpixels*views/s", "cores": 1, "kind": "reference",
                                    "sample": "whole workload; the reference has no CPU path: its CUDA build (sm_100 recompile) on 1 GPU, host side single-threaded"}
        else:
            if not args.no_parity:
                line["parity_bits_equal"] = parity_check(case, res["crc"], np)
                line["parity_note"] = "planes/states/views of the last timed configuration vs one untimed run of the reference's CUDA build (oracle/_ref) on the same inputs and seed, CRC32 of the raw bytes"
    del case
    torch.cuda.empty_cache()
    # ---- secondary block: BASELINE configs[1] (all-STRONG FIRST_INIT), N = 1 only
    if rank == 0 and world == 1 and not args.no_secondary and args.workload == "cfg3":
        sec_args = argparse.Namespace(**vars(args))
        if args.impl == "ours":
            r2, c2, _ = run_ours("cfg2", sec_args, 0, 1, local, torch, np, dist, full=False)
        else:
            r2, c2, _ = run_reference("cfg2", sec_args, local, torch, np, full=False)
        w2 = WORKLOADS["cfg2"]
        sec = {"workload": "cfg2: " + w2["desc"], "value": round(w2["W"] * w2["H"] * w2["S"] / (r2["iter_ms"] * 1e-3) / 1e6, 2), "unit": "Mpixels*views/s",
               "iter_ms": round(r2["iter_ms"], 3), "ms_per_step": round(r2["dev_step_ms"], 3), "steps": r2["steps"],
               "kernel_ms": {k: round(v, 3) for k, v in r2["kernel_ms"].items()}, "output_crc32": r2["crc"]}
        if args.impl == "ours" and not args.no_parity:
            sec["parity_bits_equal"] = parity_check(c2, r2["crc"], np)
        line["secondary"] = {"cfg2": sec}
        del c2
    # ---- secondary block: BASELINE configs[3] through the sharded pass scheduler, every N (collective: all ranks take part)
    if args.impl == "ours" and not args.no_secondary and args.workload == "cfg3":
        c4 = run_cfg4(rank, world, local, torch, np, dist)
        if rank == 0:
            line.setdefault("secondary", {})["cfg4"] = c4
    if rank == 0:
        if args.impl == "ours" and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(np, T, E)
        print(json.dumps(line))
    if world > 1:
        if args.impl == "ours":
            dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(np, T, E):
    """CPU port (oracle/apd_cpu.c) timed on a bounded sample: one colour pass of the strong propagation over a
    crop of a 1280x960, 9-source-view scene, all host threads (OpenMP)."""
    try:
        from oracle import cpu_binding as CB
        import golden_tools as G
        CB.lib()
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": "Mpixels*views/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}
    W, H, S = 1280, 960, 9
    case = T.build_case(W, H, S, iters=1, device="cpu")
    p = G.oracle_params(case)
    st, _ = CB.run(case["images"], case["cameras"], p, stage_end=4)      # K1 + K5 state
    x0, y0, x1, y1 = 40, 40, 1240, 920
    t0 = time.perf_counter()
    n = CB.strong_pass(case["images"], case["cameras"], p, st, 0, 0, x0, y0, x1, y1)
    dt = time.perf_counter() - t0
    # one colour pass = half an iteration: Mpixels*views/s per iteration = (2n pixels * S) / (2 dt)
    return {"value": round(n * S / dt / 1e6, 4), "unit": "Mpixels*views/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"one colour of the strong propagation (K6) over a {x1 - x0}x{y1 - y0} crop ({n} pixels) of a {W}x{H}, {S}-source-view scene, {dt:.1f} s wall on {os.cpu_count()} threads (K5 state prepared before, untimed)"}


if __name__ == "__main__":
    main()
