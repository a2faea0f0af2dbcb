"""The reference's on-disk formats (APD.cpp:3-92, main.cpp:6-49) through the std-only C++ of libapd_b200.so
(include/apd_io.h): .dmb/.bin matrices, *_cam.txt, pair.txt."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import engine as E
from .scene import CAMERA_DTYPE

CV_8UC1, CV_32SC1, CV_32FC1, CV_32FC3 = 0, 4, 5, 21
_DTYPES = {CV_8UC1: (np.uint8, 1), CV_32SC1: (np.int32, 1), CV_32FC1: (np.float32, 1), CV_32FC3: (np.float32, 3)}


def _lib():
    L = E.lib()
    if not getattr(L, "_io_bound", False):
        vp, ci, cs = C.c_void_p, C.c_int, C.c_char_p
        L.apd_io_elem_size.argtypes = [ci]; L.apd_io_elem_size.restype = C.c_size_t
        L.apd_io_read_mat_header.argtypes = [cs, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
        L.apd_io_read_mat.argtypes = [cs, vp, C.c_size_t, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
        L.apd_io_write_mat.argtypes = [cs, vp, ci, ci, ci]
        L.apd_io_read_camera.argtypes = [cs, vp]
        L.apd_io_read_pairs.argtypes = [cs, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), ci, ci]
        L.apd_io_format_index.argtypes = [ci, C.c_char_p]; L.apd_io_format_index.restype = None
        L._io_bound = True
    return L


def _ck(rc, what):
    if rc:
        raise E.ApdError(f"{what}: libapd_b200 io error {rc}")


def ReadBinMat(path: str) -> np.ndarray:
    L = _lib()
    r, c, t = C.c_int(), C.c_int(), C.c_int()
    _ck(L.apd_io_read_mat_header(str(path).encode(), C.byref(r), C.byref(c), C.byref(t)), path)
    dt, ch = _DTYPES[t.value]
    out = np.empty((r.value, c.value) + ((ch,) if ch > 1 else ()), dtype=dt)
    _ck(L.apd_io_read_mat(str(path).encode(), C.c_void_p(out.ctypes.data), out.nbytes, C.byref(r), C.byref(c), C.byref(t)), path)
    return out


def WriteBinMat(path: str, mat: np.ndarray):
    mat = np.ascontiguousarray(mat)
    if mat.dtype == np.uint8 and mat.ndim == 2: t = CV_8UC1
    elif mat.dtype in (np.int32, np.uint32) and mat.ndim == 2: t = CV_32SC1
    elif mat.dtype == np.float32 and mat.ndim == 2: t = CV_32FC1
    elif mat.dtype == np.float32 and mat.ndim == 3 and mat.shape[2] == 3: t = CV_32FC3
    else:
        raise ValueError(f"no reference matrix type for {mat.dtype} {mat.shape}")
    _ck(_lib().apd_io_write_mat(str(path).encode(), C.c_void_p(mat.ctypes.data), mat.shape[0], mat.shape[1], t), path)


def ReadCamera(path: str) -> np.ndarray:
    cam = np.zeros(1, dtype=CAMERA_DTYPE)
    _ck(_lib().apd_io_read_camera(str(path).encode(), C.c_void_p(cam.ctypes.data)), path)
    return cam[0]


def GenerateSampleList(pair_path: str):
    """[(ref_image_id, [src_image_ids...])] in file order (main.cpp:6-49)."""
    L = _lib()
    n = C.c_int()
    _ck(L.apd_io_read_pairs(str(pair_path).encode(), C.byref(n), None, None, None, 0, 0), pair_path)
    cap = 64
    refs = (C.c_int * max(n.value, 1))(); cnt = (C.c_int * max(n.value, 1))(); src = (C.c_int * (max(n.value, 1) * cap))()
    _ck(L.apd_io_read_pairs(str(pair_path).encode(), C.byref(n), refs, cnt, src, n.value, cap), pair_path)
    return [(refs[i], [src[i * cap + j] for j in range(cnt[i])]) for i in range(n.value)]


def ToFormatIndex(index: int) -> str:
    buf = C.create_string_buffer(9)
    _lib().apd_io_format_index(index, buf)
    return buf.value.decode()
