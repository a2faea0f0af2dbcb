"""ctypes binding of oracle/_ref/libapd_cpu.so (oracle/apd_cpu.c, the plain-C CPU restatement).
TEST INFRASTRUCTURE ONLY — see the header of apd_cpu.c."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libapd_cpu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, ci = C.c_void_p, C.c_int
        L.apd_cpu_run.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, C.c_ulonglong, ci, vp, vp, vp, vp, vp, vp]
        L.apd_cpu_run.restype = ci
        L.apd_cpu_run_apd.argtypes = L.apd_cpu_run.argtypes + [vp, vp, vp, vp]
        L.apd_cpu_run_apd.restype = ci
        L.apd_cpu_strong_pass.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci]
        L.apd_cpu_strong_pass.restype = C.c_long
        _lib = L
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


class CpuState:
    def __init__(self, H, W):
        self.planes = np.zeros((H, W, 4), np.float32); self.costs = np.zeros((H, W), np.float32)
        self.views = np.zeros((H, W), np.uint32); self.states = np.zeros((H, W), np.uint8)
        self.view_weights = np.zeros((H, W, 32), np.uint8); self.rng = np.zeros((H, W, 6), np.uint32)

    def as_dict(self):
        return {k: getattr(self, k) for k in ("planes", "costs", "views", "states", "view_weights", "rng")}


def run(images, cameras, params, depths=None, planes=None, views=None, states=None, seed=1234567, stage_end=-1):
    """params: apd_mvs_b200.engine.PatchMatchParams with depth_min/max already set (APD.cpp:454-455)."""
    images = np.ascontiguousarray(images, np.float32)
    N, H, W = images.shape
    cams = np.ascontiguousarray(cameras)
    depths = None if depths is None else np.ascontiguousarray(depths, np.float32)
    planes = None if planes is None else np.ascontiguousarray(planes, np.float32)
    views = None if views is None else np.ascontiguousarray(views, np.uint32)
    states = None if states is None else np.ascontiguousarray(states, np.uint8)
    st = CpuState(H, W)
    n = lib().apd_cpu_run(W, H, N, _p(images), _p(depths), _p(cams), C.byref(params), _p(planes), _p(views), _p(states),
                          seed, stage_end, _p(st.planes), _p(st.costs), _p(st.views), _p(st.states), _p(st.view_weights), _p(st.rng))
    return st, n


def run_apd(images, cameras, params, depths=None, planes=None, views=None, states=None, seed=1234567, stage_end=-1):
    """Like run(), and also returns the deformation-path buffers: anchors [H,W,9,2] int16 ((-1,-1) = absent; valid for
    pixels that were WEAK when K3 ran), nearest [H,W,2] int16, reliable [H,W] uint8, fit_planes [H,W,4] float32."""
    images = np.ascontiguousarray(images, np.float32)
    N, H, W = images.shape
    cams = np.ascontiguousarray(cameras)
    depths = None if depths is None else np.ascontiguousarray(depths, np.float32)
    planes = None if planes is None else np.ascontiguousarray(planes, np.float32)
    views = None if views is None else np.ascontiguousarray(views, np.uint32)
    states = None if states is None else np.ascontiguousarray(states, np.uint8)
    st = CpuState(H, W)
    extra = {"anchors": np.zeros((H, W, 9, 2), np.int16), "nearest": np.zeros((H, W, 2), np.int16),
             "reliable": np.zeros((H, W), np.uint8), "fit_planes": np.zeros((H, W, 4), np.float32)}
    n = lib().apd_cpu_run_apd(W, H, N, _p(images), _p(depths), _p(cams), C.byref(params), _p(planes), _p(views), _p(states),
                              seed, stage_end, _p(st.planes), _p(st.costs), _p(st.views), _p(st.states), _p(st.view_weights), _p(st.rng),
                              _p(extra["anchors"]), _p(extra["nearest"]), _p(extra["reliable"]), _p(extra["fit_planes"]))
    return st, extra, n


def strong_pass(images, cameras, params, st: CpuState, iter_, color, x0, y0, x1, y1):
    images = np.ascontiguousarray(images, np.float32)
    N, H, W = images.shape
    cams = np.ascontiguousarray(cameras)
    return lib().apd_cpu_strong_pass(W, H, N, _p(images), _p(cams), C.byref(params), _p(st.planes), _p(st.costs), _p(st.views),
                                     _p(st.states), _p(st.view_weights), _p(st.rng), iter_, color, x0, y0, x1, y1)
