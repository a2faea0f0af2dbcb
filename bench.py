#!/usr/bin/env python
"""bench.py — headline benchmark of the PatchMatch hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|mid]

metric   Mpixels*views/s per PatchMatch iteration = W*H*S / t_iter / 1e6, t_iter = GPU time of one
         iteration of the loop APD.cu:2443-2457 (strong black + strong red [+ fit + weak black + weak
         red]), averaged over the iterations of all timed steps.
step     one full RunPatchMatch (all launches of APD.cu:2409-2471, `iters` PatchMatch iterations) on one
         reference view of the workload, inputs already resident in HBM.
e2e      the same metric through the C-ABI with HOST buffers: every step uploads the image stack and the
         cameras from pinned host memory, runs, and reads planes/states/views back (what ProcessProblem,
         main.cpp:95-124, does per view); t_iter_e2e = step time / iters.
N > 1    reference views sharded one per rank (weak scaling, no data-path collective); rank 0 renders the
         view ring and broadcasts images+cameras once over NCCL; value = all ranks' units / max time.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (W, H, n_src, iters, description)
    "cfg2": (3111, 2074, 9, 3, "synthetic ETH3D-half-res shape 3111x2074, 1 ref x 9 src views, 3 iters (BASELINE.json configs[1])"),
    "cfg1": (256, 256, 1, 1, "synthetic 2-view 256x256, 1 iter (BASELINE.json configs[0])"),
    "mid": (1024, 768, 9, 3, "synthetic 1024x768, 1 ref x 9 src, 3 iters (development size)"),
    "cfg5": (4096, 4096, 16, 8, "synthetic 4096x4096, 1 ref x 16 src views, all STRONG, 8 iters (SURVEY §8d cfg 5: strong-kernel sweep)"),
}
TAPS_PER_PIXEL_VIEW_ITER = 14 * 36          # 14 hypotheses x 36 taps
ALG_BYTES_PER_TAP = 8                       # one fp32 reference sample + one fp32 source sample


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    W, H, S, iters, desc = WORKLOADS[args.workload]
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
    from apd_mvs_b200.scene import make_scene, CAMERA_DTYPE

    # ---- inputs: a ring of S + world views; rank r uses view r as reference and the S next views as sources
    n_views = S + (world if args.impl == "ours" else 1)
    dev = f"cuda:{local}"
    if rank == 0:
        scene = make_scene(W, H, n_views - 1, device=dev)
        images_all = scene["images"]
        cams_all = torch.from_numpy(scene["cameras"].view(np.uint8).reshape(n_views, 112).copy()).to(dev)
    else:
        images_all = torch.empty((n_views, H, W), dtype=torch.float32, device=dev)
        cams_all = torch.empty((n_views, 112), dtype=torch.uint8, device=dev)
    setup_bcast_ms = 0.0
    if world > 1 and args.impl == "ours":
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        dist.broadcast(images_all, 0); dist.broadcast(cams_all, 0)     # the single setup collective
        torch.cuda.synchronize()
        setup_bcast_ms = 1e3 * (time.perf_counter() - t0)
    order = [(rank + k) % n_views for k in range(S + 1)]
    images_dev = images_all[order].contiguous()
    cams = cams_all[order].cpu().numpy().copy().view(CAMERA_DTYPE).reshape(-1)
    del images_all
    params = E.default_params(max_iterations=iters, state=E.FIRST_INIT, use_APD=0, geom_consistency=0)
    seed = 1234567 + rank
    names = T.stage_names(iters)
    iter_idx = [i for i, n in enumerate(names) if n.startswith("it")]
    strong_idx = [i for i, n in enumerate(names) if "strong" in n]
    npx = W * H
    sampler = ClockSampler(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    result = {}
    if args.impl == "ours":
        images_host = images_dev.cpu().pin_memory()
        pb = E.Problem(images_dev, cams, params, seed=seed, device=local)
        apd = E.APD(pb)
        apd.InuputInitialization(); apd.CudaSpaceInitialization(); apd.SetDataPassHelperInCuda()
        for _ in range(args.warmup):
            apd.RunPatchMatch()
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_ms = np.zeros(len(names))
        t0 = time.perf_counter()
        step_ms = []
        for _ in range(args.steps):
            s0 = time.perf_counter()
            apd.RunPatchMatch()                      # device-resident inputs; events on the engine's own stream
            step_ms.append(1e3 * (time.perf_counter() - s0))
            stage_ms += apd.StageMs()
        barrier()
        total_ms = 1e3 * (time.perf_counter() - t0)
        clocks = sampler.stop()
        launches = apd.LaunchCount() * args.steps
        dev_step_ms = float(stage_ms.sum() / args.steps)
        iter_ms = float(stage_ms[iter_idx].sum() / (args.steps * iters))
        strong_ms = float(stage_ms[strong_idx].sum() / (args.steps * len(strong_idx)))
        # ---- e2e through the C-ABI with host buffers
        out_planes = torch.empty((H, W, 4), dtype=torch.float32).pin_memory().numpy()
        out_states = torch.empty((H, W), dtype=torch.uint8).pin_memory().numpy()
        out_views = torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        L = E.lib()
        import ctypes as C
        ptrs = (C.c_void_p * (S + 1))(*[images_host.data_ptr() + i * W * H * 4 for i in range(S + 1)])
        cam_arr = np.ascontiguousarray(cams)

        def e2e_step():
            L.apd_set_cameras(apd._h, C.c_void_p(cam_arr.ctypes.data))
            L.apd_set_images(apd._h, ptrs, W * 4)
            L.apd_run(apd._h)
            L.apd_get_planes(apd._h, C.c_void_p(out_planes.ctypes.data))
            L.apd_get_states(apd._h, C.c_void_p(out_states.ctypes.data))
            L.apd_get_views(apd._h, C.c_void_p(out_views.ctypes.data))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
        result.update(dev_step_ms=dev_step_ms, iter_ms=iter_ms, strong_ms=strong_ms, total_ms=total_ms, e2e_ms=e2e_ms,
                      launches=launches, stage_ms=(stage_ms / args.steps).round(3).tolist(), clocks=clocks,
                      h2d=(S + 1) * npx * 4 + 112 * (S + 1), d2h=npx * 21)
        apd.close()
    else:
        from oracle import ref_binding
        if not ref_binding.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libapd_ref.so not built (needs /root/reference at build time)"}))
            return
        images_np = images_dev.cpu().numpy()
        case = {"images": images_np, "cameras": cams, "params": params, "depths": None, "planes": None, "views": None, "states": None}
        stage_acc = None
        e2e_list = []
        sampler_started = False
        for it in range(args.warmup + args.steps):
            if it == args.warmup:
                torch.cuda.synchronize(); sampler.start(); sampler_started = True; t_all = time.perf_counter()
            s0 = time.perf_counter()
            ref = T.make_reference(case, seed=seed)          # construct + upload  (APD.cpp:356-699)
            ref.run()                                        # APD::RunPatchMatch (incl. its own D2H, APD.cu:2490-2492)
            ref.outputs()
            ms = ref.stage_ms()
            ref.close()                                      # ~APD
            if it >= args.warmup:
                e2e_list.append(1e3 * (time.perf_counter() - s0))
                stage_acc = ms if stage_acc is None else stage_acc + ms
        total_ms = 1e3 * (time.perf_counter() - t_all)
        clocks = sampler.stop() if sampler_started else {}
        stage_ms = stage_acc / args.steps
        result.update(dev_step_ms=float(stage_ms.sum()), iter_ms=float(stage_ms[iter_idx].sum() / iters),
                      strong_ms=float(stage_ms[strong_idx].mean()), total_ms=total_ms, e2e_ms=float(np.mean(e2e_list)),
                      launches=25 * args.steps, stage_ms=np.round(stage_ms[:len(names)], 3).tolist(), clocks=clocks, h2d=0, d2h=0)

    # ---- max over ranks
    vals = torch.tensor([result["iter_ms"], result["dev_step_ms"], result["e2e_ms"], result["total_ms"], result["strong_ms"]],
                        dtype=torch.float64, device=dev)
    if world > 1 and args.impl == "ours":
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    iter_ms, dev_step_ms, e2e_ms, total_ms, strong_ms = [float(v) for v in vals.tolist()]
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
        # dominant kernel = one colour of the strong propagation: (npx/2) pixels x S views x 504 taps x 8 B
        traffic = None
        try:   # per-launch dram__bytes_read+write of the dominant kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if args.impl == "ours" and tj.get("workload") == args.workload:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        alg_bytes = (npx / 2) * S * TAPS_PER_PIXEL_VIEW_ITER * ALG_BYTES_PER_TAP
        achieved = alg_bytes / (strong_ms * 1e-3) / 1e9
        line = {
            "metric": "Mpixels*views/s per PatchMatch iteration", "value": round(value, 2), "unit": "Mpixels*views/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_step_ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": args.impl,
            "config": {"workload": f"{args.workload}: {desc}", "width": W, "height": H, "src_views": S, "iters": iters,
                       "ref_views_per_step": n_units, "state": "FIRST_INIT", "l2": "inputs+state larger than L2 (126 MB)" if npx * (4 * (S + 1) + 130) > 126e6 else "fits L2; state rewritten every step",
                       "parallelism": f"ref-views-dp{world}"},
            "iter_ms": round(iter_ms, 3), "strong_kernel_ms": round(strong_ms, 3), "wall_ms_per_step": round(total_ms / args.steps, 3),
            "stage_ms": result["stage_ms"],
            "e2e": {"value": round(e2e_value, 2), "unit": "Mpixels*views/s", "ms_per_call": round(e2e_ms, 3),
                    "h2d_bytes_per_step": result["h2d"], "d2h_bytes_per_step": result["d2h"]},
            "gpu_launches": result["launches"],
            "clocks": result["clocks"],
            "roofline": {"bound": "hbm", "kernel": "k_strong (K6/K7)" if args.impl == "ours" else "Black/RedPixelUpdateStrong",
                         "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                         "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "note": "algorithmic bytes = 8 B per NCC tap x 504 taps per pixel*view of the REFERENCE algorithm (SURVEY §8d), independent of what the kernel chooses to fetch: exact skips (zero-weight views, hypotheses that can no longer win) can push frac above 1; the kernel is bound by the texture/LSU pipes and issue slots, DRAM traffic is ~3% of this (profiles/)"},
        }
        if world > 1:
            line["setup_broadcast_ms"] = round(setup_bcast_ms, 2)
        if args.impl == "reference":
            line["cpu_baseline"] = {"value": round(value, 2), "unit": "Mpixels*views/s", "cores": 1, "kind": "reference",
                                    "sample": "whole workload; the reference has no CPU path: its CUDA build (sm_100 recompile) on 1 GPU, host side single-threaded"}
            line["e2e"]["h2d_bytes_per_step"] = 0; line["e2e"]["d2h_bytes_per_step"] = 0
        elif world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(np, T, E)
        print(json.dumps(line))
    if world > 1:
        if args.impl == "ours":
            dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(np, T, E):
    """CPU port (oracle/apd_cpu.c) timed on a bounded sample: one colour pass of the strong propagation over a
    crop of a 640x480, 9-source-view scene, all host threads (OpenMP)."""
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
