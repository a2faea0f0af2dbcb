"""ctypes binding of oracle/_ref/libapd_ref.so — the reference's own APD.cu recompiled for sm_100
(see oracle/ref_wrapper.cu). TEST / BASELINE INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's reference arm, never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libapd_ref.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, ci = C.c_void_p, C.c_int
        L.apdref_create.argtypes = [C.POINTER(vp), ci, ci, ci, ci, vp, C.c_ulonglong]
        L.apdref_upload.argtypes = [vp, C.POINTER(vp), C.c_size_t, vp, C.POINTER(vp), C.c_size_t, vp, vp, vp]
        L.apdref_want_snapshots.argtypes = [vp, C.POINTER(ci), ci]
        L.apdref_run.argtypes = [vp, ci]
        L.apdref_get_stage_ms.argtypes = [vp, C.POINTER(C.c_double), ci]
        L.apdref_get.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
        L.apdref_get_outputs.argtypes = [vp, vp, vp, vp]
        L.apdref_get_anchors.argtypes = [vp, vp, vp, vp, vp, vp]
        L.apdref_last_error.argtypes = [vp]; L.apdref_last_error.restype = C.c_char_p
        L.apdref_destroy.argtypes = [vp]; L.apdref_destroy.restype = None
        _lib = L
    return _lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


class RefAPD:
    """Drives the unmodified reference APD::RunPatchMatch() on caller-supplied arrays."""

    def __init__(self, images, cameras, params, depths=None, planes=None, views=None, states=None,
                 seed: int = 1234567, device: int = 0):
        self.L = lib()
        self.images = np.ascontiguousarray(images, dtype=np.float32)
        self.N, self.H, self.W = self.images.shape
        self.h = C.c_void_p(None)
        params.num_images = self.N
        rc = self.L.apdref_create(C.byref(self.h), device, self.W, self.H, self.N, C.byref(params), seed)
        if rc:
            raise RuntimeError(f"apdref_create {rc}")
        cams = np.ascontiguousarray(cameras)
        iptr = (C.c_void_p * self.N)(*[self.images.ctypes.data + i * self.W * self.H * 4 for i in range(self.N)])
        dptr = None
        if depths is not None:
            self.depths = np.ascontiguousarray(depths, dtype=np.float32)
            dptr = (C.c_void_p * self.N)(*[self.depths.ctypes.data + i * self.W * self.H * 4 for i in range(self.N)])
        planes = None if planes is None else np.ascontiguousarray(planes, dtype=np.float32)
        views = None if views is None else np.ascontiguousarray(views, dtype=np.uint32)
        states = None if states is None else np.ascontiguousarray(states, dtype=np.uint8)
        self._ck(self.L.apdref_upload(self.h, iptr, self.W * 4, _ptr(cams), dptr, self.W * 4,
                                      _ptr(planes), _ptr(views), _ptr(states)))

    def _ck(self, rc):
        if rc < 0:
            raise RuntimeError(f"libapd_ref error {rc}: {self.L.apdref_last_error(self.h).decode()}")

    def run(self, snapshots=(), quiet=True):
        arr = (C.c_int * len(snapshots))(*snapshots)
        self.L.apdref_want_snapshots(self.h, arr, len(snapshots))
        self._ck(self.L.apdref_run(self.h, 1 if quiet else 0))

    def stage_ms(self):
        buf = (C.c_double * 256)()
        n = self.L.apdref_get_stage_ms(self.h, buf, 256)
        return np.array(buf[:n])

    def get(self, stage=-1):
        H, W = self.H, self.W
        out = {"planes": np.empty((H, W, 4), np.float32), "costs": np.empty((H, W), np.float32),
               "views": np.empty((H, W), np.uint32), "states": np.empty((H, W), np.uint8),
               "view_weights": np.empty((H, W, 32), np.uint8), "rng": np.empty((H, W, 12), np.uint32)}
        self._ck(self.L.apdref_get(self.h, stage, _ptr(out["planes"]), _ptr(out["costs"]), _ptr(out["views"]),
                                   _ptr(out["states"]), _ptr(out["view_weights"]), _ptr(out["rng"])))
        # curandStateXORWOW: d, v[5], boxmuller_flag, boxmuller_flag_double, float, double -> (v0..v4, d)
        r = out["rng"]
        out["rng"] = np.concatenate([r[..., 1:6], r[..., 0:1]], axis=-1)
        return out

    def outputs(self):
        H, W = self.H, self.W
        planes = np.empty((H, W, 4), np.float32); states = np.empty((H, W), np.uint8); views = np.empty((H, W), np.uint32)
        self.L.apdref_get_outputs(self.h, _ptr(planes), _ptr(states), _ptr(views))
        return planes, states, views

    def anchors(self):
        """Returns dense [H,W,9,2] anchors ((-1,-1) where absent / pixel not WEAK at upload time)."""
        H, W = self.H, self.W
        nmap = np.empty((H, W), np.int32); nearest = np.empty((H, W, 2), np.int16)
        reliable = np.empty((H, W), np.uint8); fit = np.empty((H, W, 4), np.float32)
        wc = self.L.apdref_get_anchors(self.h, None, _ptr(nmap), _ptr(nearest), _ptr(reliable), _ptr(fit))
        comp = np.empty((max(wc, 1), 9, 2), np.int16)
        self._ck(self.L.apdref_get_anchors(self.h, _ptr(comp), None, None, None, None))
        return comp, nmap, nearest, reliable, fit, wc

    def close(self):
        if self.h:
            self.L.apdref_destroy(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
