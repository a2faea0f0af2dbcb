#!/usr/bin/env python
"""Fusion on the GPU (include/apd_fusion.h) against the reference's unmodified RunFusion on the host CPU
(oracle/_ref/libapd_fusion_ref.so), on the depth maps of a real run of the pass schedule.

    python tests/tools/fusion_bench.py --width 1920 --height 1080 --views 8 --src 5 [--out profiles/...json]"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--src", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import fusion_tools as FT
    from apd_mvs_b200 import fusion as F, pipeline as P
    from apd_mvs_b200.scene import make_scene

    W, H, V, S = args.width, args.height, args.views, args.src
    sc = make_scene(W, H, V - 1, device="cuda")
    images, cams = sc["images"].cpu().numpy(), sc["cameras"]
    pairs = P.ring_pairs(V, S)
    scene = P.Scene(images, cams, pairs, seed=5)
    scene.Run()
    depths = [scene.Depth(v) for v in range(V)]; normals = [scene.Normal(v) for v in range(V)]; states = [scene.States(v) for v in range(V)]
    scene.close()
    bgr = FT.colour_images(images)
    fu = F.Fusion(V, W, H)
    t0 = time.perf_counter()
    for v in range(V):
        fu.SetView(v, bgr[v], cams[v], depths[v], normals[v], states[v])
    for r, s in pairs:
        fu.AddProblem(r, s)
    t1 = time.perf_counter()
    xyz, col = fu.RunFusion()
    t2 = time.perf_counter()
    tm = fu.Timing()
    line = {"workload": f"{V} views {W}x{H}, {S} source views each", "points": int(len(xyz)),
            "ours": {"gpu_ms": round(tm["gpu_ms"], 2), "run_wall_ms_incl_point_download": round(1e3 * (t2 - t1), 1), "upload_ms": round(1e3 * (t1 - t0), 1),
                     "max_decision_rounds_per_view": tm["max_rounds"]}}
    if FT.ref_available():
        with tempfile.TemporaryDirectory() as d:
            FT.write_dense_folder(d, list(range(V)), bgr, cams, depths, normals, states)
            t0 = time.perf_counter()
            rxyz, rbgr = FT.run_reference_fusion(d, [(r, s) for r, s in pairs])
            ref_ms = 1e3 * (time.perf_counter() - t0)
        same = len(xyz) == len(rxyz) and np.array_equal(xyz.view(np.uint32), rxyz.view(np.uint32)) and np.array_equal(col.astype(np.uint8), rbgr)
        diff = 0
        if not same:
            a = {x.tobytes() for x in xyz}; b = {x.tobytes() for x in rxyz}
            diff = len(a ^ b)
        line["reference"] = {"cpu_ms_incl_file_io": round(ref_ms, 1), "points": int(len(rxyz)), "cores": 1}
        line["bit_identical_point_list"] = bool(same)
        line["points_not_in_common"] = diff
        line["speedup_vs_reference_cpu"] = round(ref_ms / (1e3 * (t2 - t1)), 1)
    fu.close()
    print(json.dumps(line))
    if args.out:
        json.dump(line, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
