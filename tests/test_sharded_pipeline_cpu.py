"""world_size-2 gloo test of the sharded pass schedule (apd_mvs_b200.pipeline.ShardedScene): problem
ownership, per-pass depth-map broadcasts and the visibility rule (own results of this pass, peers' results of the
previous pass) against a single-process emulation of the same rule. The PatchMatch is replaced by a cheap
deterministic stand-in that depends on every depth map it is allowed to see."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_VIEWS, N_SRC, ROUNDS = 5, 2, 2
SIZES = [(6, 4), (12, 8)]


def pairs():
    return [(r, [(r + k) % N_VIEWS for k in range(1, N_SRC + 1)]) for r in range(N_VIEWS)]


class FakeBackend:
    """Stand-in for pipeline.Scene: depth maps are CPU tensors in full-size buffers, like the device buffers."""

    def __init__(self):
        self.buf = [torch.zeros(SIZES[-1][0] * SIZES[-1][1]) for _ in range(N_VIEWS)]
        self.size = [None] * N_VIEWS
        self.pairs = pairs()

    def round_size(self, i): return SIZES[i]
    def sync(self): pass
    def depth_tensor(self, v, w, h): return self.buf[v][: w * h].view(h, w)
    def mark_result(self, v, w, h): self.size[v] = (w, h)

    def visible(self, v, w, h):
        if self.size[v] is None:
            return torch.zeros(h, w)
        sw, sh = self.size[v]
        d = self.depth_tensor(v, sw, sh)
        return d if (sw, sh) == (w, h) else d.repeat_interleave(2, 0).repeat_interleave(2, 1)[:h, :w]

    def process(self, i, ps, k):
        ref, srcs = self.pairs[k]
        w, h = SIZES[i]
        acc = 0.25 * self.visible(ref, w, h).clone()
        if ps > 0:                                        # geometric passes read the sources' depth maps
            for s_ in srcs:
                acc += 0.5 * self.visible(s_, w, h)
        acc += (i * 4 + ps + 1) + 0.01 * k + torch.arange(w * h, dtype=torch.float32).view(h, w) * 1e-3
        self.depth_tensor(ref, w, h).copy_(acc)
        self.size[ref] = (w, h)


def emulate(world):
    b = FakeBackend()
    for i in range(ROUNDS):
        for ps in range(4):
            start_buf = [x.clone() for x in b.buf]; start_size = list(b.size)
            merged_buf = [x.clone() for x in b.buf]; merged_size = list(b.size)
            for r in range(world):
                b.buf = [x.clone() for x in start_buf]; b.size = list(start_size)
                for k in range(r, N_VIEWS, world):
                    b.process(i, ps, k)
                    ref = b.pairs[k][0]
                    merged_buf[ref] = b.buf[ref].clone(); merged_size[ref] = b.size[ref]
            b.buf, b.size = merged_buf, merged_size
    return b


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from apd_mvs_b200.pipeline import ShardedScene
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = FakeBackend()
    sh = ShardedScene(b, pairs(), rank, world, ROUNDS)
    mine = sh.my_problems()
    sh.run()
    q.put((rank, mine, [x.numpy().copy() for x in b.buf], list(b.size)))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_schedule_matches_emulation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    want = emulate(2)
    for rank, _, bufs, sizes in res:
        assert sizes == [SIZES[-1]] * N_VIEWS
        for v in range(N_VIEWS):
            assert np.array_equal(bufs[v], want.buf[v].numpy()), (rank, v)
    # and the sharded result differs from the sequential (Gauss-Seidel) one, as SURVEY §3.1 says it must
    seq = emulate(1)
    assert any(not np.array_equal(seq.buf[v].numpy(), want.buf[v].numpy()) for v in range(N_VIEWS))


def test_single_rank_is_the_reference_order():
    sys.path.insert(0, ROOT)
    from apd_mvs_b200.pipeline import ShardedScene
    b = FakeBackend()
    ShardedScene(b, pairs(), 0, 1, ROUNDS).run()
    seq = emulate(1)
    for v in range(N_VIEWS):
        assert np.array_equal(b.buf[v].numpy(), seq.buf[v].numpy())
