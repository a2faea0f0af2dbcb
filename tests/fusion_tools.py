"""Test helpers for the fusion: write a dense_folder the UNMODIFIED reference RunFusion can read (oracle/_ref/
libapd_fusion_ref.so, built from /root/reference/APD.cpp against oracle/shim_host), run it, read its PLY."""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libapd_fusion_ref.so")


def ref_available():
    return os.path.exists(REF_LIB)


def fmt(i):
    return "%08d" % i


def write_mat(path, arr):
    """ReadBinMat layout, APD.cpp:3-29 (independent of the product's writer)."""
    arr = np.ascontiguousarray(arr)
    code = {(np.dtype(np.uint8), 2): 0, (np.dtype(np.int32), 2): 4, (np.dtype(np.float32), 2): 5, (np.dtype(np.float32), 3): 21}[(arr.dtype, arr.ndim)]
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", 1, arr.shape[0], arr.shape[1], code)); f.write(arr.tobytes())


def write_dense_folder(root, ids, bgr, cams, depths, normals, weaks):
    """images/%08d.jpg hold the raw container the oracle's imread shim parses (not JPEG)."""
    os.makedirs(os.path.join(root, "images"), exist_ok=True); os.makedirs(os.path.join(root, "cams"), exist_ok=True)
    os.makedirs(os.path.join(root, "APD"), exist_ok=True)
    for k, i in enumerate(ids):
        img = np.ascontiguousarray(bgr[k], np.uint8)
        with open(os.path.join(root, "images", fmt(i) + ".jpg"), "wb") as f:
            f.write(b"APDRAW\0\0" + struct.pack("<3i", img.shape[0], img.shape[1], 3) + img.tobytes())
        c = cams[k]
        R, t, K = c["R"].reshape(3, 3), c["t"], c["K"].reshape(3, 3)
        g = lambda x: "%.9g" % float(x)          # 9 significant digits round-trip a float32
        txt = "extrinsic\n" + "".join(" ".join(g(x) for x in list(R[r]) + [t[r]]) + "\n" for r in range(3)) + "0.0 0.0 0.0 1.0\n\nintrinsic\n"
        txt += "".join(" ".join(g(x) for x in K[r]) + "\n" for r in range(3)) + f"\n{g(c['depth_min'])} 0.01 192 {g(c['depth_max'])}\n"
        open(os.path.join(root, "cams", fmt(i) + "_cam.txt"), "w").write(txt)
        out = os.path.join(root, "APD", fmt(i)); os.makedirs(out, exist_ok=True)
        write_mat(os.path.join(out, "depths.dmb"), depths[k].astype(np.float32))
        write_mat(os.path.join(out, "normals.dmb"), normals[k].astype(np.float32))
        write_mat(os.path.join(out, "weak.bin"), weaks[k].astype(np.uint8))


def run_reference_fusion(root, problems, variant=0):
    """problems: [(ref_image_id, [src_image_ids])]. Returns (xyz float32 [n,3], bgr uint8 [n,3]) read from APD/APD.ply.
    variant 0 = RunFusion, 1 = RunFusion_TAT_Intermediate, 2 = RunFusion_TAT_advanced."""
    L = C.CDLL(REF_LIB)
    L.apdfusion_ref_run.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    L.apdfusion_ref_run_variant.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int]
    n = len(problems); ms = max(1, max(len(s) for _, s in problems))
    refs = (C.c_int * n)(*[r for r, _ in problems]); cnt = (C.c_int * n)(*[len(s) for _, s in problems])
    src = (C.c_int * (n * ms))()
    for i, (_, s) in enumerate(problems):
        for j, v in enumerate(s):
            src[i * ms + j] = v
    assert L.apdfusion_ref_run_variant(str(root).encode(), n, refs, cnt, src, ms, variant) == 0
    return read_ply(os.path.join(root, "APD", "APD.ply"))


def read_ply(path):
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode()
    n = int([l for l in header.splitlines() if l.startswith("element vertex")][0].split()[-1])
    rec = np.dtype([("xyz", "<f4", 3), ("bgr", "u1", 3)])
    a = np.frombuffer(raw, dtype=rec, count=n, offset=end)
    return a["xyz"].copy(), a["bgr"].copy()


def colour_images(gray):
    """Three different channels from a grey stack [V,H,W] -> uint8 [V,H,W,3] (B, G, R)."""
    g = np.clip(gray, 0, 255)
    return np.stack([g, 255.0 - g, 0.5 * g + 30.0], axis=-1).astype(np.uint8)
