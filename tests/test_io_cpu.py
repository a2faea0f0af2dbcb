"""CPU tests of the std-only file formats (include/apd_io.h) against byte layouts written by hand from the
reference's ReadBinMat / WriteBinMat (APD.cpp:3-50), ReadCamera (APD.cpp:52-92) and GenerateSampleList (main.cpp:6-49)."""
import struct

import numpy as np
import pytest

from apd_mvs_b200 import engine as E
from apd_mvs_b200 import io as IO


def test_write_mat_byte_layout(tmp_path):
    d = np.arange(6, dtype=np.float32).reshape(2, 3) * 0.5
    p = tmp_path / "depths.dmb"
    IO.WriteBinMat(p, d)
    raw = p.read_bytes()
    assert raw[:16] == struct.pack("<4i", 1, 2, 3, 5)            # version, rows, cols, CV_32FC1
    assert raw[16:] == d.tobytes()


@pytest.mark.parametrize("arr,code", [
    (np.random.default_rng(0).random((5, 7)).astype(np.float32), 5),
    (np.random.default_rng(1).random((4, 6, 3)).astype(np.float32), 21),
    (np.random.default_rng(2).integers(0, 3, (9, 2)).astype(np.uint8), 0),
    (np.random.default_rng(3).integers(0, 2 ** 31 - 1, (3, 3)).astype(np.int32), 4),
])
def test_mat_round_trip(tmp_path, arr, code):
    p = tmp_path / "m.bin"
    IO.WriteBinMat(p, arr)
    assert struct.unpack("<4i", p.read_bytes()[:16]) == (1, arr.shape[0], arr.shape[1], code)
    back = IO.ReadBinMat(p)
    assert back.dtype == arr.dtype and np.array_equal(back, arr)


def test_selected_views_are_stored_as_32sc1(tmp_path):
    v = np.array([[0xFFFFFFFF, 5]], dtype=np.uint32)              # CV_32SC1 holding unsigned masks (APD.cpp:551)
    p = tmp_path / "selected_views.bin"
    IO.WriteBinMat(p, v)
    assert np.array_equal(IO.ReadBinMat(p).view(np.uint32), v)


def test_read_hand_written_mat(tmp_path):
    p = tmp_path / "weak.bin"
    p.write_bytes(struct.pack("<4i", 1, 2, 2, 0) + bytes([0, 1, 2, 1]))
    assert np.array_equal(IO.ReadBinMat(p), np.array([[0, 1], [2, 1]], np.uint8))


def test_bad_version_and_missing_file(tmp_path):
    p = tmp_path / "bad.dmb"
    p.write_bytes(struct.pack("<4i", 2, 1, 1, 5) + b"\0\0\0\0")
    with pytest.raises(E.ApdError):
        IO.ReadBinMat(p)
    with pytest.raises(E.ApdError):
        IO.ReadBinMat(tmp_path / "nope.dmb")


CAM = """extrinsic
0.970263 0.00747983 0.241939 -191.02
-0.0147429 0.999493 0.0282234 3.28832
-0.241605 -0.030951 0.969881 22.5401
0.0 0.0 0.0 1.0

intrinsic
2892.33 0 823.205
0 2883.18 619.071
0 0 1

425.0 2.5 192 935.0
"""


def test_read_camera(tmp_path):
    p = tmp_path / "00000000_cam.txt"
    p.write_text(CAM)
    cam = IO.ReadCamera(p)
    R = np.array([0.970263, 0.00747983, 0.241939, -0.0147429, 0.999493, 0.0282234, -0.241605, -0.030951, 0.969881], np.float32)
    t = np.array([-191.02, 3.28832, 22.5401], np.float32)
    assert np.array_equal(cam["R"], R) and np.array_equal(cam["t"], t)
    assert np.array_equal(cam["K"], np.array([2892.33, 0, 823.205, 0, 2883.18, 619.071, 0, 0, 1], np.float32))
    c = -(R.reshape(3, 3).astype(np.float64).T @ t.astype(np.float64))
    # the reference accumulates R[0+j]*t0 + R[3+j]*t1 + R[6+j]*t2 in double, left to right, then rounds to float
    want = np.array([-np.float32(float(R[0 + j]) * float(t[0]) + float(R[3 + j]) * float(t[1]) + float(R[6 + j]) * float(t[2])) for j in range(3)], np.float32)
    assert np.array_equal(cam["c"], want) and np.allclose(cam["c"], c, rtol=1e-6)
    assert cam["depth_min"] == np.float32(425.0) and cam["depth_max"] == np.float32(935.0)


PAIRS = """3
0
3 1 100.5 2 50.25 7 0.0
1
2 0 80.0 2 -1.0
2
2 1 3.5 0 2.5
"""


def test_generate_sample_list_drops_non_positive_scores(tmp_path):
    p = tmp_path / "pair.txt"
    p.write_text(PAIRS)
    assert IO.GenerateSampleList(p) == [(0, [1, 2]), (1, [0]), (2, [1, 0])]


def test_format_index():
    assert IO.ToFormatIndex(7) == "00000007" and IO.ToFormatIndex(12345678) == "12345678"


# ---- pinned against the reference's OWN file code (APD.cpp:3-92, compiled unmodified into oracle/_ref/libapd_fusion_ref.so)
def _ref_io():
    import ctypes as C
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libapd_fusion_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libapd_fusion_ref.so not built")
    L = C.CDLL(path)
    if not hasattr(L, "apdref_write_bin_mat"):
        pytest.skip("oracle/_ref/libapd_fusion_ref.so predates the file-I/O wrappers")
    L.apdref_write_bin_mat.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.apdref_read_bin_mat.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.apdref_read_camera.argtypes = [C.c_char_p, C.c_void_p]
    return L, C


@pytest.mark.parametrize("arr,code", [
    (np.random.default_rng(10).random((37, 53)).astype(np.float32) * 9.0, 5),            # depths.dmb   CV_32FC1
    (np.random.default_rng(11).standard_normal((37, 53, 3)).astype(np.float32), 21),      # normals.dmb  CV_32FC3
    (np.random.default_rng(12).integers(0, 3, (37, 53)).astype(np.uint8), 0),             # weak.bin     CV_8UC1
    (np.random.default_rng(13).integers(0, 2 ** 31 - 1, (37, 53)).astype(np.int32), 4),   # selected_views.bin CV_32SC1
])
def test_files_written_by_the_reference_are_read_back_and_vice_versa(tmp_path, arr, code):
    L, C = _ref_io()
    a = np.ascontiguousarray(arr)
    # the reference's WriteBinMat -> this library's reader
    p = tmp_path / "ref_written.dmb"
    assert L.apdref_write_bin_mat(str(p).encode(), a.ctypes.data, a.shape[0], a.shape[1], code) == 0
    back = IO.ReadBinMat(p)
    assert back.dtype == a.dtype and back.shape == a.shape and np.array_equal(back.view(np.uint8), a.view(np.uint8))
    # this library's writer -> the reference's ReadBinMat; and both writers produce the same bytes
    q = tmp_path / "ours_written.dmb"
    IO.WriteBinMat(q, a)
    assert q.read_bytes() == p.read_bytes()
    out = np.empty_like(a)
    r, c, t = C.c_int(), C.c_int(), C.c_int()
    assert L.apdref_read_bin_mat(str(q).encode(), out.ctypes.data, out.nbytes, C.byref(r), C.byref(c), C.byref(t)) == 0
    assert (r.value, c.value, t.value) == (a.shape[0], a.shape[1], code) and np.array_equal(out.view(np.uint8), a.view(np.uint8))


def test_camera_file_parsed_like_the_reference(tmp_path):
    L, C = _ref_io()
    p = tmp_path / "00000003_cam.txt"
    p.write_text(CAM)
    mine = IO.ReadCamera(p)
    ref = np.zeros(1, dtype=mine.dtype)
    assert L.apdref_read_camera(str(p).encode(), ref.ctypes.data) == 0
    for f in ("K", "R", "t", "c"):
        assert np.array_equal(np.asarray(mine[f]).view(np.uint32), np.asarray(ref[0][f]).view(np.uint32)), f
    assert mine["depth_min"] == ref[0]["depth_min"] and mine["depth_max"] == ref[0]["depth_max"]
