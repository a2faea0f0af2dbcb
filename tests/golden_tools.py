"""Load a golden fixture (tests/golden/*.npz, made by tests/golden/make_golden.py from the reference)."""
from __future__ import annotations
import ctypes as C
import os
import numpy as np
from apd_mvs_b200 import engine as E
from apd_mvs_b200.scene import CAMERA_DTYPE

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    p = E.PatchMatchParams()
    C.memmove(C.byref(p), g["params"].tobytes(), C.sizeof(p))
    cams = np.ascontiguousarray(g["cameras"]).view(CAMERA_DTYPE).reshape(-1)
    case = {"params": p, "images": g["images"], "cameras": cams, "depths": g.get("in_depths"), "planes": g.get("in_planes"),
            "views": g.get("in_views"), "states": g.get("in_states")}
    return g, case


def oracle_params(case):
    """Params as the reference host fills them (APD.cpp:454-457)."""
    p = E.PatchMatchParams()
    C.memmove(C.byref(p), C.byref(case["params"]), C.sizeof(p))
    cams = case["cameras"]
    p.depth_min = float(np.float32(cams[0]["depth_min"]) * np.float32(0.6))
    p.depth_max = float(np.float32(cams[0]["depth_max"]) * np.float32(1.2))
    p.num_images = len(cams)
    return p
