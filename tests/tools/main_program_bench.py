#!/usr/bin/env python
"""Wall time of the reference's own program (oracle/_ref/apd_main_ref) and of the same main.cpp linked against the facade +
libapd_b200.so (oracle/_ref/apd_main_b200) on one synthetic dense_folder; checks that APD.ply is byte-identical.

    python tests/tools/main_program_bench.py --width 1000 --height 750 --views 6 --src 4 [--out profiles/...json]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1000)
    ap.add_argument("--height", type=int, default=750)
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--src", type=int, default=4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import fusion_tools as FT
    import test_main_program_gpu as TM
    from apd_mvs_b200.scene import make_scene
    W, H, V, S = args.width, args.height, args.views, args.src
    assert max(W, H) <= 1000, "the oracle's OpenCV stand-in has no cv::resize: one round only"
    sc = make_scene(W, H, V - 1, device="cuda")
    bgr = FT.colour_images(sc["images"].cpu().numpy())
    pairs = [(r, [(r + k) % V for k in range(1, S + 1)]) for r in range(V)]
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "ref"), os.path.join(d, "ours")
        TM.write_inputs(a, list(range(V)), bgr, sc["cameras"], pairs); TM.write_inputs(b, list(range(V)), bgr, sc["cameras"], pairs)
        t0 = time.perf_counter(); r1 = subprocess.run([TM.REF_EXE, a, "0"], capture_output=True, text=True)
        t1 = time.perf_counter(); r2 = subprocess.run([TM.OUR_EXE, b, "0"], capture_output=True, text=True, env=dict(os.environ, APD_SEED="1234567", APD_B200_TIMING="1"))
        t2 = time.perf_counter()
        assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr[-500:], r2.stderr[-500:])
        same = open(os.path.join(a, "APD", "APD.ply"), "rb").read() == open(os.path.join(b, "APD", "APD.ply"), "rb").read()
        n = len(FT.read_ply(os.path.join(a, "APD", "APD.ply"))[0])

        def cost_ms(out):   # "Cost time: N ms" per ProcessProblem (main.cpp:137)
            return sum(int(l.split()[2]) for l in out.splitlines() if l.startswith("Cost time:"))
        line = {"workload": f"{V} views {W}x{H}, {S} source views each, 1 round x 4 passes = {4 * V} ProcessProblem calls + RunFusion (CPU)",
                "reference_program_s": round(t1 - t0, 2), "facade_build_s": round(t2 - t1, 2),
                "reference_sum_ProcessProblem_ms": cost_ms(r1.stdout), "facade_sum_ProcessProblem_ms": cost_ms(r2.stdout),
                "fused_points": n, "ply_byte_identical": bool(same),
                "facade_own_time": next((l for l in r2.stderr.splitlines() if l.startswith("[apd_b200 facade]")), None)}
        if os.environ.get("APD_B200_POOL") == "0":
            line["pool"] = "off"
    print(json.dumps(line))
    if args.out:
        json.dump(line, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
