"""Host-side mirror of the reference's RunFusion (APD.cpp:826-977) on top of the GPU fusion of libapd_b200.so
(include/apd_fusion.h). Thin ctypes calls only; no CPU fallback."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import engine as E
from .scene import CAMERA_DTYPE


def _bind(lib):
    if getattr(lib, "_fusion_bound", False):
        return lib
    vp, ci = C.c_void_p, C.c_int
    lib.apd_fusion_create.argtypes = [C.POINTER(vp), ci, ci, ci, ci]
    lib.apd_fusion_destroy.argtypes = [vp]; lib.apd_fusion_destroy.restype = None
    lib.apd_fusion_last_error.argtypes = [vp]; lib.apd_fusion_last_error.restype = C.c_char_p
    lib.apd_fusion_set_view.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
    lib.apd_fusion_set_view_planes.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
    lib.apd_fusion_add_problem.argtypes = [vp, ci, C.POINTER(ci), ci]
    lib.apd_fusion_run.argtypes = [vp]
    lib.apd_fusion_run_tat.argtypes = [vp, ci]
    lib.apd_fusion_num_points.argtypes = [vp]; lib.apd_fusion_num_points.restype = C.c_longlong
    lib.apd_fusion_get_points.argtypes = [vp, vp, vp]
    lib.apd_fusion_write_ply.argtypes = [vp, C.c_char_p]
    lib.apd_fusion_get_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(ci)]
    lib._fusion_bound = True
    return lib


class Fusion:
    def __init__(self, n_views: int, width: int, height: int, device: int = 0):
        self.L = _bind(E.lib())
        self.n_views, self.W, self.H = n_views, width, height
        self._h = C.c_void_p(None)
        rc = self.L.apd_fusion_create(C.byref(self._h), device, n_views, width, height)
        if rc:
            raise E.ApdError(f"apd_fusion_create failed ({rc})")

    def SetView(self, view, bgr, camera, depth, normal, states, block=None):
        """bgr [H,W,3] uint8, camera CAMERA_DTYPE scalar, depth [H,W] f32, normal [H,W,3] f32, states [H,W] u8."""
        bgr = np.ascontiguousarray(bgr, np.uint8); depth = np.ascontiguousarray(depth, np.float32)
        normal = np.ascontiguousarray(normal, np.float32); states = np.ascontiguousarray(states, np.uint8)
        cam = np.ascontiguousarray(np.asarray(camera, dtype=CAMERA_DTYPE).reshape(1))
        blk = None if block is None else np.ascontiguousarray(block, np.uint8)
        self._ck(self.L.apd_fusion_set_view(self._h, view, bgr.ctypes.data, cam.ctypes.data, depth.ctypes.data, normal.ctypes.data,
                                            states.ctypes.data, None if blk is None else blk.ctypes.data))

    def AddProblem(self, ref, srcs):
        arr = (C.c_int * max(len(srcs), 1))(*srcs)
        self._ck(self.L.apd_fusion_add_problem(self._h, ref, arr, len(srcs)))

    def RunFusion(self, variant: str = "eth"):
        """variant: "eth" = RunFusion (APD.cpp:826-977), "tat_intermediate" / "tat_advanced" = RunFusion_TAT_Intermediate /
        RunFusion_TAT_advanced (APD.cpp:979-1296)."""
        if variant == "eth":
            self._ck(self.L.apd_fusion_run(self._h))
        else:
            self._ck(self.L.apd_fusion_run_tat(self._h, {"tat_intermediate": 1, "tat_advanced": 2}[variant]))
        n = self.L.apd_fusion_num_points(self._h)
        xyz = np.empty((n, 3), np.float32); col = np.empty((n, 3), np.float32)
        self._ck(self.L.apd_fusion_get_points(self._h, xyz.ctypes.data, col.ctypes.data))
        return xyz, col

    def ExportPointCloud(self, path):
        self._ck(self.L.apd_fusion_write_ply(self._h, str(path).encode()))

    def Timing(self):
        ms, r = C.c_double(), C.c_int()
        self.L.apd_fusion_get_timing(self._h, C.byref(ms), C.byref(r))
        return {"gpu_ms": ms.value, "max_rounds": r.value}

    def close(self):
        if self._h:
            self.L.apd_fusion_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise E.ApdError(f"libapd_b200 fusion error {rc}: {self.L.apd_fusion_last_error(self._h).decode()}")
