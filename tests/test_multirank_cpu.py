"""world_size-2 gloo test of the N>1 host path: setup broadcast of images+cameras from rank 0, per-rank
view-ring selection, max-over-ranks timing reduction (what bench.py does around the kernels)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from apd_mvs_b200 import shard
    from apd_mvs_b200.scene import make_scene
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, n_views = 3, 3 + world
    if rank == 0:
        sc = make_scene(48, 32, n_views - 1)
        images = sc["images"].clone()
        cams = torch.from_numpy(sc["cameras"].view(np.uint8).reshape(n_views, 112).copy())
    else:
        images = torch.zeros((n_views, 32, 48)); cams = torch.zeros((n_views, 112), dtype=torch.uint8)
    shard.broadcast_inputs(images, cams, 0)
    order = shard.view_order(rank, S, n_views)
    t = shard.max_over_ranks(torch.tensor([1.0 + rank], dtype=torch.float64))
    q.put((rank, order, float(images.sum()), int(cams.sum()), float(t[0])))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_sharding_with_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    (r0, o0, s0, c0, t0), (r1, o1, s1, c1, t1) = res
    assert o0 == [0, 1, 2, 3] and o1 == [1, 2, 3, 4]
    assert s0 == s1 and c0 == c1 and s0 != 0          # both ranks hold the broadcast stack
    assert t0 == t1 == 2.0                             # max over ranks
