#!/usr/bin/env python
"""Whole-schedule timing of the scene layer (include/apd_scene.h) against the reference's per-(problem, pass)
object lifetime, on the same box and the same synthetic multi-view scene (SURVEY §8d cfg 4, scaled by --views).

    python tests/tools/pipeline_bench.py --width 1920 --height 1080 --views 8 --src 5 [--no-reference]

ours      : Scene.Run() = 4*round_num passes over all problems, everything resident in HBM.
reference : for every (problem, pass) construct + upload + APD::RunPatchMatch + download + destroy of the UNMODIFIED
            reference (oracle/_ref), i.e. ProcessProblem without its JPEG decoding and .dmb file traffic (which the
            real reference also pays); the numpy host logic of oracle/ref_pipeline.py is NOT counted.
Also checks that both pipelines end bit-identical. Prints one JSON line; --out writes it to a file."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--src", type=int, default=5)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    import parity_tools as T
    from apd_mvs_b200 import engine as E, pipeline as P
    from apd_mvs_b200.scene import make_scene
    from oracle import ref_pipeline as RP, ref_binding

    W, H, V, S = args.width, args.height, args.views, args.src
    sc = make_scene(W, H, V - 1, device="cuda")
    images, cams = sc["images"].cpu().numpy(), sc["cameras"]
    pairs = P.ring_pairs(V, S)
    scene = P.Scene(images, cams, pairs, seed=99)
    rounds = scene.ComputeRoundNum()
    scene.RunPass(0, 0)                       # warm-up (module load, first-touch), then start over
    scene.close()
    scene = P.Scene(images, cams, pairs, seed=99)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scene.Run()
    ours_wall = 1e3 * (time.perf_counter() - t0)
    tm = scene.Timing()
    line = {"workload": f"{V} views {W}x{H}, {S} source views each, {rounds} rounds x 4 passes = {4 * rounds * V} PatchMatch runs",
            "ours": {"wall_ms": round(ours_wall, 1), "patchmatch_gpu_ms": round(tm["patchmatch_ms"], 1), "launches": tm["launches"]}}
    if not args.no_reference and ref_binding.available():
        acc = {"life_ms": 0.0, "gpu_ms": 0.0, "runs": 0, "create_upload_ms": 0.0, "run_ms": 0.0, "download_ms": 0.0, "destroy_ms": 0.0}

        def run_ref(imgs, cams_, params, depths, planes, views, states, seed):
            t = time.perf_counter()
            ref = ref_binding.RefAPD(imgs, cams_, T.clone_params(params), depths=depths, planes=planes, views=views, states=states, seed=seed)
            t1 = time.perf_counter()
            ref.run()
            t2 = time.perf_counter()
            out = ref.outputs()
            t3 = time.perf_counter()
            acc["gpu_ms"] += float(ref.stage_ms().sum())
            t3b = time.perf_counter()
            ref.close()
            t4 = time.perf_counter()
            acc["create_upload_ms"] += 1e3 * (t1 - t); acc["run_ms"] += 1e3 * (t2 - t1); acc["download_ms"] += 1e3 * (t3 - t2); acc["destroy_ms"] += 1e3 * (t4 - t3b)
            acc["life_ms"] += 1e3 * (t4 - t) - 1e3 * (t3b - t3)
            acc["runs"] += 1
            return out
        rp = RP.RefPipeline(images, cams, pairs, E.default_params, run_ref, seed=99)
        t0 = time.perf_counter()
        rp.run()
        ref_total = 1e3 * (time.perf_counter() - t0)
        same = all(np.array_equal(scene.Depth(v).view(np.uint32), rp.results[v]["depth"].view(np.uint32)) and
                   np.array_equal(scene.Normal(v).view(np.uint32), rp.results[v]["normal"].view(np.uint32)) and
                   np.array_equal(scene.States(v), rp.results[v]["weak"]) and
                   np.array_equal(scene.SelectedViews(v), rp.results[v]["views"]) for v in range(V))
        line["reference"] = {"object_lifetimes_ms": round(acc["life_ms"], 1), "patchmatch_gpu_ms": round(acc["gpu_ms"], 1),
                             "runs": acc["runs"], "with_numpy_host_logic_ms": round(ref_total, 1),
                             "phases_ms": {k: round(acc[k], 1) for k in ("create_upload_ms", "run_ms", "download_ms", "destroy_ms")}}
        line["bit_identical"] = bool(same)
        line["speedup_wall_vs_object_lifetimes"] = round(acc["life_ms"] / ours_wall, 2)
        line["speedup_patchmatch_gpu"] = round(acc["gpu_ms"] / tm["patchmatch_ms"], 2)
    d = scene.Depth(0)
    gt = sc["depth"][0].cpu().numpy()
    ok = d > 0
    line["accuracy_view0"] = {"valid_frac": round(float(ok.mean()), 4),
                              "median_rel_err": float(np.median(np.abs(d[ok] - gt[ok]) / gt[ok])),
                              "frac_rel_err_lt_1e-2": round(float((np.abs(d[ok] - gt[ok]) / gt[ok] < 1e-2).mean()), 4)}
    scene.close()
    print(json.dumps(line))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(line, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
