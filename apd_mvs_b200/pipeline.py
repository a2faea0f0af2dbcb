"""Host-side mirror of the reference's driver loop (main.cpp:140-217) on top of the scene layer of
libapd_b200.so (include/apd_scene.h): all views, pyramid levels and per-view results stay on the GPU across the
4 * round_num passes. Thin ctypes calls only; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import engine as E
from .scene import CAMERA_DTYPE


def _bind(lib):
    if getattr(lib, "_scene_bound", False):
        return lib
    vp, ci = C.c_void_p, C.c_int
    lib.apd_scene_create.argtypes = [C.POINTER(vp), ci, ci, ci, ci, C.c_uint64]
    lib.apd_scene_destroy.argtypes = [vp]; lib.apd_scene_destroy.restype = None
    lib.apd_scene_last_error.argtypes = [vp]; lib.apd_scene_last_error.restype = C.c_char_p
    lib.apd_scene_set_view.argtypes = [vp, ci, vp, C.c_size_t, vp]
    lib.apd_scene_add_problem.argtypes = [vp, ci, C.POINTER(ci), ci]
    lib.apd_scene_set_round_limit.argtypes = [vp, ci]
    lib.apd_scene_num_rounds.argtypes = [vp]
    lib.apd_scene_round_size.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
    lib.apd_scene_pass_params.argtypes = [vp, ci, ci, C.POINTER(E.PatchMatchParams)]
    lib.apd_scene_run.argtypes = [vp]
    lib.apd_scene_run_pass.argtypes = [vp, ci, ci]
    lib.apd_scene_run_problem.argtypes = [vp, ci, ci, ci]
    lib.apd_scene_result_size.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
    for name in ("depth", "normal", "states", "views"):
        getattr(lib, "apd_scene_get_" + name).argtypes = [vp, ci, vp]
    lib.apd_scene_set_result.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp]
    lib.apd_scene_result_device.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.apd_scene_mark_result.argtypes = [vp, ci, ci, ci]
    lib.apd_scene_get_scaled_image.argtypes = [vp, ci, ci, vp]
    lib.apd_scene_get_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib._scene_bound = True
    return lib


class Scene:
    """One dense_folder of the reference: images + cameras of all views and the pair list."""

    def __init__(self, images, cameras, pairs, seed: int = 1234567, device: int = 0, round_limit: int | None = None):
        """images: [n_views, H, W] float32 (numpy, or a torch CUDA tensor on `device`); cameras: CAMERA_DTYPE[n_views];
        pairs: list of (ref_view, [src_views...]) in pair.txt order; round_limit: replaces the literal 1000 of
        ComputeRoundNum (main.cpp:81)."""
        self.L = _bind(E.lib())
        n, H, W = images.shape
        self.n_views, self.H, self.W = int(n), int(H), int(W)
        self.device = device
        self._h = C.c_void_p(None)
        rc = self.L.apd_scene_create(C.byref(self._h), device, self.n_views, self.W, self.H, seed)
        if rc:
            raise E.ApdError(f"apd_scene_create failed ({rc})")
        if round_limit is not None:
            self._ck(self.L.apd_scene_set_round_limit(self._h, int(round_limit)))
        cams = np.ascontiguousarray(cameras, dtype=CAMERA_DTYPE)
        for v in range(self.n_views):
            if isinstance(images, np.ndarray):
                img = np.ascontiguousarray(images[v], dtype=np.float32)
                ptr = img.ctypes.data
            else:
                img = images[v].contiguous()
                ptr = img.data_ptr()
            self._ck(self.L.apd_scene_set_view(self._h, v, C.c_void_p(ptr), W * 4, C.c_void_p(cams[v:v + 1].ctypes.data)))
        self.pairs = [(int(r), [int(x) for x in s]) for r, s in pairs]
        for r, s in self.pairs:
            arr = (C.c_int * len(s))(*s)
            self._ck(self.L.apd_scene_add_problem(self._h, r, arr, len(s)))

    # ---- main.cpp vocabulary
    def ComputeRoundNum(self) -> int:  # main.cpp:72-88
        return self.L.apd_scene_num_rounds(self._h)

    def RoundSize(self, round_: int):
        w, h = C.c_int(), C.c_int()
        self._ck(self.L.apd_scene_round_size(self._h, round_, C.byref(w), C.byref(h)))
        return w.value, h.value

    def PassParams(self, round_: int, pass_: int) -> E.PatchMatchParams:  # main.cpp:171-211
        p = E.PatchMatchParams()
        self._ck(self.L.apd_scene_pass_params(self._h, round_, pass_, C.byref(p)))
        return p

    def Run(self):  # main.cpp:168-217
        self._ck(self.L.apd_scene_run(self._h))

    def RunPass(self, round_: int, pass_: int):
        self._ck(self.L.apd_scene_run_pass(self._h, round_, pass_))

    def ProcessProblem(self, round_: int, pass_: int, problem: int):  # main.cpp:91-138
        self._ck(self.L.apd_scene_run_problem(self._h, round_, pass_, problem))

    # ---- results (depths.dmb, normals.dmb, weak.bin, selected_views.bin of a view)
    def ResultSize(self, view: int):
        w, h = C.c_int(), C.c_int()
        self._ck(self.L.apd_scene_result_size(self._h, view, C.byref(w), C.byref(h)))
        return w.value, h.value

    def _get(self, name, view, shape_tail, dtype):
        w, h = self.ResultSize(view)
        out = np.empty((h, w) + shape_tail, dtype=dtype)
        self._ck(getattr(self.L, "apd_scene_get_" + name)(self._h, view, C.c_void_p(out.ctypes.data)))
        return out

    def Depth(self, view): return self._get("depth", view, (), np.float32)
    def Normal(self, view): return self._get("normal", view, (3,), np.float32)
    def States(self, view): return self._get("states", view, (), np.uint8)
    def SelectedViews(self, view): return self._get("views", view, (), np.uint32)

    def SetResult(self, view: int, depth, normal, states, selected_views):
        """Checkpoint / resume: load a view's maps (as returned by Depth / Normal / States / SelectedViews or read from
        its depths.dmb / normals.dmb / weak.bin / selected_views.bin)."""
        d = np.ascontiguousarray(depth, np.float32); n = np.ascontiguousarray(normal, np.float32)
        st = np.ascontiguousarray(states, np.uint8); sv = np.ascontiguousarray(selected_views).view(np.uint32)
        h, w = d.shape
        self._ck(self.L.apd_scene_set_result(self._h, view, w, h, d.ctypes.data, n.ctypes.data, st.ctypes.data, sv.ctypes.data))

    # ---- multi-GPU hand-over (used by ShardedScene)
    def depth_tensor(self, view: int, width: int, height: int):
        """torch CUDA tensor aliasing the view's device depth buffer, shaped for a width x height result."""
        import torch
        d = C.c_void_p()
        self._ck(self.L.apd_scene_result_device(self._h, view, None, C.byref(d), None, None))

        class _Alias:
            __cuda_array_interface__ = {"shape": (height, width), "typestr": "<f4", "data": (d.value, False), "version": 2}
        return torch.as_tensor(_Alias(), device=f"cuda:{self.device}")

    def mark_result(self, view: int, width: int, height: int):
        self._ck(self.L.apd_scene_mark_result(self._h, view, width, height))

    def round_size(self, round_: int):
        return self.RoundSize(round_)

    def sync(self):
        import torch
        torch.cuda.synchronize(self.device)

    def process(self, round_: int, pass_: int, problem: int):
        self.ProcessProblem(round_, pass_, problem)

    def ScaledImage(self, round_: int, view: int):
        w, h = self.RoundSize(round_)
        out = np.empty((h, w), np.float32)
        self._ck(self.L.apd_scene_get_scaled_image(self._h, round_, view, C.c_void_p(out.ctypes.data)))
        return out

    def Timing(self):
        pm, wall, n = C.c_double(), C.c_double(), C.c_longlong()
        self.L.apd_scene_get_timing(self._h, C.byref(pm), C.byref(wall), C.byref(n))
        return {"patchmatch_ms": pm.value, "wall_ms": wall.value, "launches": n.value}

    def close(self):
        if self._h:
            self.L.apd_scene_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise E.ApdError(f"libapd_b200 scene error {rc}: {self.L.apd_scene_last_error(self._h).decode()}")


def ring_pairs(n_views: int, n_src: int):
    """pair.txt of a ring of views: every view is a reference, its sources are the next n_src views."""
    return [(r, [(r + k) % n_views for k in range(1, n_src + 1)]) for r in range(n_views)]


class ShardedScene:
    """The schedule of main.cpp:168-217 over `world` ranks: problem k belongs to rank k % world; every
    rank holds all images and cameras (one broadcast at setup) and, after each pass, ONE all-gather hands every rank the
    depth maps the others produced - the only data another rank's next pass reads (geometric term, APD.cpp:492-510).
    Priors (normals, pixel states, selected views) never leave their owner. Within a rank problems keep pair-list order,
    so a rank sees its own earlier results of the same pass (as the reference does) and its peers' results of the
    previous pass: the sharded schedule is block-Jacobi where the reference is Gauss-Seidel, which is why
    parity is defined per (problem, pass) on identical inputs (the tests emulate this visibility rule around the
    reference).

    `backend` needs process(round, pass, k), depth_tensor(view, w, h), mark_result(view, w, h), round_size(round), sync();
    `Scene` is the GPU backend, the CPU tests use a numpy stand-in."""

    def __init__(self, backend, pairs, rank: int, world: int, rounds: int):
        self.b, self.pairs, self.rank, self.world, self.rounds = backend, pairs, rank, world, rounds
        self.process_ms = 0.0
        self.exchange_ms = 0.0
        self.wait_ms = 0.0
        refs = [r for r, _ in pairs]
        if len(set(refs)) != len(refs):
            raise ValueError("a view may be the reference of one problem only (pair.txt has one entry per image)")
        self._stage = {}

    def owner(self, problem: int) -> int:
        return problem % self.world

    def my_problems(self):
        return [k for k in range(len(self.pairs)) if self.owner(k) == self.rank]

    def run_pass(self, round_: int, pass_: int):
        import time
        t0 = time.perf_counter()
        for k in self.my_problems():
            self.b.process(round_, pass_, k)
        self.b.sync()
        t1 = time.perf_counter()
        self.process_ms += 1e3 * (t1 - t0)
        self.exchange(round_)

    def exchange(self, round_: int):
        """One all-gather per pass: every rank packs the depth maps it owns into a contiguous stack (slot j = its j-th
        problem; ranks with one problem less pad the last slot), the gathered [world, slots, h, w] stack is unpacked into
        the peers' result buffers in place."""
        if self.world == 1:
            return
        import time
        import torch
        import torch.distributed as dist
        w, h = self.b.round_size(round_)
        n = len(self.pairs)
        slots = (n + self.world - 1) // self.world
        mine = self.my_problems()
        ref0 = self.b.depth_tensor(self.pairs[0][0], w, h)
        key = (w, h)
        if key not in self._stage:
            self._stage = {key: (torch.zeros((slots, h, w), dtype=ref0.dtype, device=ref0.device),
                                 torch.empty((self.world, slots, h, w), dtype=ref0.dtype, device=ref0.device))}
        stage, gathered = self._stage[key]
        for j, k in enumerate(mine):
            stage[j].copy_(self.b.depth_tensor(self.pairs[k][0], w, h))
        self.b.sync()
        t0 = time.perf_counter()
        dist.barrier()                              # waiting for the slowest owner, reported apart from the transfer
        t1 = time.perf_counter()
        dist.all_gather(list(gathered.unbind(0)), stage)
        for k, (ref, _) in enumerate(self.pairs):
            r = self.owner(k)
            if r != self.rank:
                self.b.depth_tensor(ref, w, h).copy_(gathered[r, k // self.world])
                self.b.mark_result(ref, w, h)
        self.b.sync()        # the scene's own stream must not run ahead of the collective's stream
        self.wait_ms += 1e3 * (t1 - t0)
        self.exchange_ms += 1e3 * (time.perf_counter() - t1)

    def run(self, passes=range(4)):
        for i in range(self.rounds):
            for pass_ in passes:
                self.run_pass(i, pass_)
