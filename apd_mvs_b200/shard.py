"""Reference-view sharding across ranks: one unit = one (reference view, pass) =
one RunPatchMatch call; units share only read-only inputs, so ranks need no data-path collective.
Rank 0 broadcasts the image stack + cameras once at setup; timing is the max over ranks."""
from __future__ import annotations


def view_order(unit: int, n_src: int, n_views: int):
    """Indices into the shared view ring for work unit `unit`: [reference, src_1 .. src_S]."""
    if n_src + 1 > n_views:
        raise ValueError("ring too small")
    return [(unit + k) % n_views for k in range(n_src + 1)]


def units_of_rank(rank: int, world: int, n_units: int):
    """Round-robin assignment of reference views to ranks."""
    return list(range(rank, n_units, world))


def broadcast_inputs(images, cameras, src: int = 0):
    """The single setup collective: image stack and camera table from `src` to every rank."""
    import torch.distributed as dist
    dist.broadcast(images, src)
    dist.broadcast(cameras, src)


def max_over_ranks(t):
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
