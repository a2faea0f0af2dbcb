"""Host-side mirror of the reference's `class APD` (APD.h:67-145) on top of libapd_b200.so.

Python is only the test/bench host here: every call below is a thin ctypes call into the C-ABI of
include/apd_b200.h. The method names are the reference's own (including the `InuputInitialization`
typo, APD.h:72) so that a ProcessProblem-style driver (main.cpp:91-138) reads the same. There is NO
CPU fallback: if the CUDA library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from .scene import CAMERA_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libapd_b200.so")

FIRST_INIT, REFINE_INIT, REFINE_ITER = 0, 1, 2
WEAK, STRONG, UNKNOWN = 0, 1, 2
MAX_IMAGES = 32


class PatchMatchParams(C.Structure):
    """struct PatchMatchParams, main.h:75-94 (72 bytes)."""
    _fields_ = [("max_iterations", C.c_int), ("num_images", C.c_int), ("sigma_spatial", C.c_float),
                ("sigma_color", C.c_float), ("top_k", C.c_int), ("depth_min", C.c_float), ("depth_max", C.c_float),
                ("geom_consistency", C.c_ubyte), ("_pad0", C.c_ubyte * 3),
                ("strong_radius", C.c_int), ("strong_increment", C.c_int), ("weak_radius", C.c_int),
                ("weak_increment", C.c_int), ("use_APD", C.c_ubyte), ("_pad1", C.c_ubyte * 3),
                ("weak_peak_radius", C.c_int), ("rotate_time", C.c_int), ("ransac_threshold", C.c_float),
                ("geom_factor", C.c_float), ("state", C.c_int)]


assert C.sizeof(PatchMatchParams) == 72


def default_params(**kw) -> PatchMatchParams:
    p = PatchMatchParams(3, 5, 5.0, 3.0, 4, 0.0, 1.0, 0, (C.c_ubyte * 3)(), 5, 2, 5, 5, 1, (C.c_ubyte * 3)(),
                         2, 4, 0.005, 0.2, FIRST_INIT)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class ApdError(RuntimeError):
    pass


def _load(path: str) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA library is the product; there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    lib.apd_create.argtypes = [C.POINTER(vp), ci, ci, ci, ci, C.POINTER(PatchMatchParams), C.c_uint64]
    lib.apd_destroy.argtypes = [vp]; lib.apd_destroy.restype = None
    lib.apd_last_error.argtypes = [vp]; lib.apd_last_error.restype = C.c_char_p
    lib.apd_set_params.argtypes = [vp, C.POINTER(PatchMatchParams)]
    lib.apd_set_seed.argtypes = [vp, C.c_uint64]
    lib.apd_reset_inputs.argtypes = [vp]
    lib.apd_get_capacity.argtypes = [vp]
    lib.apd_set_num_images.argtypes = [vp, ci]
    lib.apd_set_upload_mode.argtypes = [vp, ci]
    lib.apd_set_cameras.argtypes = [vp, vp]
    lib.apd_set_images.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    lib.apd_set_images_device.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    lib.apd_set_depths.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    lib.apd_set_depths_device.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    lib.apd_set_priors.argtypes = [vp, vp, vp, vp]
    lib.apd_run.argtypes = [vp]
    lib.apd_run_until.argtypes = [vp, ci]
    lib.apd_num_stages.argtypes = [vp]
    for name in ("planes", "states", "views", "costs", "view_weights", "rng"):
        getattr(lib, "apd_get_" + name).argtypes = [vp, vp]
    lib.apd_get_anchors.argtypes = [vp, vp, vp, vp, vp]
    lib.apd_get_stage_ms.argtypes = [vp, C.POINTER(cf), ci]
    lib.apd_get_launch_count.argtypes = [vp]
    lib.apd_get_stream.argtypes = [vp]; lib.apd_get_stream.restype = vp
    lib.apd_default_params.argtypes = [C.POINTER(PatchMatchParams)]; lib.apd_default_params.restype = None
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load(LIB_PATH)
    return _lib


def _ptr(a) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


class Problem:
    """The part of `struct Problem` (main.h:96-106) that reaches the hot path, with arrays instead of
    folders: images [N,H,W] float32, cameras CAMERA_DTYPE[N] (index 0 = reference view)."""

    def __init__(self, images, cameras, params: PatchMatchParams, depths=None, planes=None, views=None,
                 states=None, seed: int = 1234567, device: int = 0):
        self.images, self.cameras, self.params = images, cameras, params
        self.depths, self.planes, self.views, self.states = depths, planes, views, states
        self.seed, self.device = seed, device


class APD:
    """Drop-in for the reference's `class APD` public surface (APD.h:69-81)."""

    def __init__(self, problem: Problem):
        self.problem = problem
        self._h = C.c_void_p(None)
        self._L = lib()
        self.width = self.height = self.num_images = 0

    # -- reference API -----------------------------------------------------------------------------
    def InuputInitialization(self):  # sic, APD.h:72 / APD.cpp:399-583
        pb = self.problem
        imgs = pb.images
        self.num_images, self.height, self.width = imgs.shape
        if self.num_images > MAX_IMAGES:
            raise ApdError(f"Can't process so much images: {self.num_images}")  # APD.cpp:428-431
        cams = np.ascontiguousarray(pb.cameras, dtype=CAMERA_DTYPE)
        pb.params.num_images = self.num_images
        pb.params.depth_min = float(np.float32(cams[0]["depth_min"]) * np.float32(0.6))   # APD.cpp:454
        pb.params.depth_max = float(np.float32(cams[0]["depth_max"]) * np.float32(1.2))   # APD.cpp:455
        self._cams = cams

    def CudaSpaceInitialization(self):  # APD.cpp:585-671
        pb, L = self.problem, self._L
        self._ck(L.apd_create(C.byref(self._h), pb.device, self.width, self.height, self.num_images,
                              C.byref(pb.params), pb.seed), create=True)
        self._ck(L.apd_set_cameras(self._h, _ptr(self._cams)))
        self._set_stack(L.apd_set_images, L.apd_set_images_device, pb.images)
        if pb.params.geom_consistency:
            if pb.depths is None:
                raise ApdError("geom_consistency needs depths (APD.cpp:492-510)")
            self._set_stack(L.apd_set_depths, L.apd_set_depths_device, pb.depths)
        if pb.planes is not None or pb.states is not None:
            planes = None if pb.planes is None else np.ascontiguousarray(pb.planes, dtype=np.float32)
            views = None if pb.views is None else np.ascontiguousarray(pb.views, dtype=np.uint32)
            states = None if pb.states is None else np.ascontiguousarray(pb.states, dtype=np.uint8)
            self._ck(L.apd_set_priors(self._h, _ptr(planes), _ptr(views), _ptr(states)))

    def SetDataPassHelperInCuda(self):  # APD.cpp:673-699: nothing left to do, arguments travel by value
        pass

    def RunPatchMatch(self, stage_end: int = -1):  # APD.cu:2386-2495
        self._ck(self._L.apd_run_until(self._h, stage_end))

    def GetPlaneHypothesis(self, r: int, c: int):  # APD.cpp:701-703
        if getattr(self, "_planes_host", None) is None:
            self._planes_host = self.GetPlaneHypotheses()
        return tuple(self._planes_host[r, c])

    def GetPixelStates(self):  # APD.cpp:705-707, CV_8UC1
        return self._get("states", (self.height, self.width), np.uint8)

    def GetSelectedViews(self):  # APD.cpp:709-711, CV_32SC1 holding unsigned bitmasks
        return self._get("views", (self.height, self.width), np.uint32)

    def GetWidth(self): return self.width
    def GetHeight(self): return self.height
    def GetDepthMin(self): return self.problem.params.depth_min
    def GetDepthMax(self): return self.problem.params.depth_max

    # -- extras -------------------------------------------------------------------------
    def GetPlaneHypotheses(self):
        return self._get("planes", (self.height, self.width, 4), np.float32)

    def GetCosts(self):
        return self._get("costs", (self.height, self.width), np.float32)

    def GetViewWeights(self):
        return self._get("view_weights", (self.height, self.width, 32), np.uint8)

    def GetRng(self):
        return self._get("rng", (self.height, self.width, 6), np.uint32)

    def GetAnchors(self):
        n = self.height * self.width
        anchors = np.empty((self.height, self.width, 9, 2), np.int16)
        nearest = np.empty((self.height, self.width, 2), np.int16)
        reliable = np.empty((self.height, self.width), np.uint8)
        fit = np.empty((self.height, self.width, 4), np.float32)
        self._ck(self._L.apd_get_anchors(self._h, _ptr(anchors), _ptr(nearest), _ptr(reliable), _ptr(fit)))
        return anchors, nearest, reliable, fit

    def SetParams(self, params: PatchMatchParams):
        self.problem.params = params
        self._ck(self._L.apd_set_params(self._h, C.byref(params)))

    def StageMs(self):
        n = self._L.apd_num_stages(self._h)
        buf = (C.c_float * n)()
        self._L.apd_get_stage_ms(self._h, buf, n)
        return np.array(buf[:], dtype=np.float64)

    def LaunchCount(self) -> int:
        return int(self._L.apd_get_launch_count(self._h))

    def NumStages(self) -> int:
        return int(self._L.apd_num_stages(self._h))

    def close(self):
        if self._h:
            self._L.apd_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------------------
    def _set_stack(self, fn_host, fn_dev, stack):
        try:
            import torch
            is_t = isinstance(stack, torch.Tensor)
        except ImportError:  # pragma: no cover
            is_t = False
        H, W = self.height, self.width
        if is_t and stack.is_cuda:
            t = stack.contiguous()
            self._ck(fn_dev(self._h, C.c_void_p(t.data_ptr()), W * 4, W * H * 4))
            return
        if is_t:
            t = stack.contiguous()
            base = t.data_ptr()
            ptrs = (C.c_void_p * self.num_images)(*[base + i * W * H * 4 for i in range(self.num_images)])
            self._ck(fn_host(self._h, ptrs, W * 4))
            self._keep = t
            return
        a = np.ascontiguousarray(stack, dtype=np.float32)
        ptrs = (C.c_void_p * self.num_images)(*[a.ctypes.data + i * W * H * 4 for i in range(self.num_images)])
        self._ck(fn_host(self._h, ptrs, W * 4))

    def _get(self, what, shape, dtype, out=None):
        a = np.empty(shape, dtype) if out is None else out
        self._ck(getattr(self._L, "apd_get_" + what)(self._h, _ptr(a)))
        return a

    def _ck(self, rc, create=False):
        if rc != 0:
            msg = self._L.apd_last_error(self._h) if self._h else b"apd_create failed"
            raise ApdError(f"libapd_b200 error {rc}: {msg.decode() if msg else ''}")


def ProcessProblem(problem: Problem):
    """main.cpp:91-124 without the disk I/O: returns depth [H,W], normal [H,W,3], states, views."""
    apd = APD(problem)
    apd.InuputInitialization()
    apd.CudaSpaceInitialization()
    apd.SetDataPassHelperInCuda()
    apd.RunPatchMatch()
    planes = apd.GetPlaneHypotheses()
    states = apd.GetPixelStates()
    depth = planes[..., 3].copy()
    bad = (depth < apd.GetDepthMin()) | (depth > apd.GetDepthMax())      # main.cpp:109-112
    depth[bad] = 0
    states[bad] = UNKNOWN
    views = apd.GetSelectedViews()
    apd.close()
    return depth, planes[..., :3].copy(), states, views
