#!/usr/bin/env python
"""Sharded pass schedule on N GPUs (one process per GPU, torchrun): problems round-robin over ranks, one NCCL
broadcast of images + cameras at setup, per-pass NCCL broadcasts of the depth maps (cfg 4 shape).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \\
        tests/tools/pipeline_sharded.py --width 1920 --height 1080 --views 8 --src 5 [--check]

--check: rank 0 also runs the reference driver emulation with the same visibility rule
(oracle.ref_pipeline.RefPipeline.run_pass(world=N)) and every rank compares the views it owns bit for bit."""
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
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from apd_mvs_b200 import engine as E, pipeline as P
    from apd_mvs_b200.scene import make_scene, CAMERA_DTYPE

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    W, H, V, S = args.width, args.height, args.views, args.src
    dev = f"cuda:{local}"
    if rank == 0:
        sc = make_scene(W, H, V - 1, device=dev)
        images = sc["images"]
        cams_t = torch.from_numpy(sc["cameras"].view(np.uint8).reshape(V, 112).copy()).to(dev)
    else:
        images = torch.empty((V, H, W), dtype=torch.float32, device=dev)
        cams_t = torch.empty((V, 112), dtype=torch.uint8, device=dev)
    if world > 1:
        dist.broadcast(images, 0); dist.broadcast(cams_t, 0)       # the setup collective
    cams = cams_t.cpu().numpy().copy().view(CAMERA_DTYPE).reshape(-1)
    pairs = P.ring_pairs(V, S)
    scene = P.Scene(images, cams, pairs, seed=99, device=local)
    rounds = scene.ComputeRoundNum()
    sh = P.ShardedScene(scene, pairs, rank, world, rounds)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sh.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = torch.tensor([1e3 * (time.perf_counter() - t0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    per_rank = torch.zeros((world, 2), dtype=torch.float64, device=dev)
    per_rank[rank, 0], per_rank[rank, 1] = sh.process_ms, sh.exchange_ms
    if world > 1:
        dist.all_reduce(per_rank)
    ok = torch.tensor([1], device=dev)
    if args.check:
        import parity_tools as T
        from oracle import ref_pipeline as RP, ref_binding
        # every rank runs the emulation on its own GPU (simple, and no result shipping is needed)
        def run_ref(imgs, cams_, params, depths, planes, views, states, seed):
            ref = ref_binding.RefAPD(imgs, cams_, T.clone_params(params), depths=depths, planes=planes, views=views,
                                     states=states, seed=seed, device=local)
            ref.run(); out = ref.outputs(); ref.close()
            return out
        rp = RP.RefPipeline(images.cpu().numpy(), cams, pairs, E.default_params, run_ref, seed=99)
        for i in range(rp.rounds):
            for ps in range(4):
                rp.run_pass(i, ps, world=world)
        good = True
        for k in sh.my_problems():
            v = pairs[k][0]
            r = rp.results[v]
            good &= np.array_equal(scene.Depth(v).view(np.uint32), r["depth"].view(np.uint32))
            good &= np.array_equal(scene.Normal(v).view(np.uint32), r["normal"].view(np.uint32))
            good &= np.array_equal(scene.States(v), r["weak"]) and np.array_equal(scene.SelectedViews(v), r["views"])
        for v in range(V):   # depth maps received from peers
            good &= np.array_equal(scene.Depth(v).view(np.uint32), rp.results[v]["depth"].view(np.uint32))
        ok[0] = 1 if good else 0
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        line = {"workload": f"{V} views {W}x{H}, {S} source views each, {rounds} rounds x 4 passes", "n_gpus": world,
                "wall_ms_max_over_ranks": round(float(wall[0]), 1), "runs_total": 4 * rounds * V,
                "runs_per_s": round(4 * rounds * V / (float(wall[0]) * 1e-3), 2),
                "per_rank_process_ms": [round(float(x), 1) for x in per_rank[:, 0]],
                "per_rank_exchange_and_wait_ms": [round(float(x), 1) for x in per_rank[:, 1]]}
        if args.check:
            line["bit_identical_to_reference_emulation"] = bool(int(ok[0]))
        print(json.dumps(line))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            json.dump(line, open(args.out, "w"), indent=1)
    scene.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
