"""CPU-only: the C-ABI library loads and exports every symbol include/apd_b200.h declares; struct
layouts match the reference's (main.h); argument errors are reported, not exit()ed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from apd_mvs_b200 import engine as E
from apd_mvs_b200.scene import CAMERA_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("apd_b200.h", "apd_scene.h", "apd_io.h", "apd_fusion.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(apd_[a-z_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported():
    lib = C.CDLL(E.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 44
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported"


def test_struct_layouts_match_reference():
    assert C.sizeof(E.PatchMatchParams) == 72          # main.h:75-94
    assert CAMERA_DTYPE.itemsize == 112                 # main.h:47-56
    assert CAMERA_DTYPE.fields["height"][1] == 96 and CAMERA_DTYPE.fields["depth_max"][1] == 108
    assert E.PatchMatchParams.geom_consistency.offset == 28 and E.PatchMatchParams.state.offset == 68


def test_default_params_are_the_reference_defaults():
    p = E.PatchMatchParams()
    E.lib().apd_default_params(C.byref(p))
    assert (p.max_iterations, p.top_k, p.strong_radius, p.strong_increment, p.weak_radius, p.weak_increment) == (3, 4, 5, 2, 5, 5)
    assert (p.use_APD, p.weak_peak_radius, p.rotate_time, p.geom_consistency) == (1, 2, 4, 0)
    assert abs(p.ransac_threshold - 0.005) < 1e-9 and abs(p.geom_factor - 0.2) < 1e-7


def test_argument_errors_are_codes_not_exits():
    L = E.lib()
    h = C.c_void_p(None)
    p = E.default_params()
    assert L.apd_create(C.byref(h), 0, 64, 64, 33, C.byref(p), 1) == -4          # > MAX_IMAGES (APD.cpp:428-431)
    assert L.apd_create(C.byref(h), 0, 40000, 64, 3, C.byref(p), 1) == -4        # short2 anchors (APD.h:48)
    p.strong_radius = 7
    assert L.apd_create(C.byref(h), 0, 64, 64, 3, C.byref(p), 1) == -4
    assert L.apd_create(None, 0, 64, 64, 3, C.byref(p), 1) == -1
    assert L.apd_run(None) == -1 and L.apd_get_planes(None, None) == -1
    assert L.apd_last_error(None) == b"null handle"


def test_no_cpu_fallback(monkeypatch, tmp_path):
    """The product must fail loudly when the CUDA library is missing."""
    with pytest.raises(ImportError):
        E._load(str(tmp_path / "libapd_b200.so"))


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "apd_mvs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                s = open(os.path.join(root, f)).read()
                assert "oracle" not in s.replace("reference oracle", ""), f"{f} mentions oracle/"


def test_oracle_libraries_export_their_entry_points():
    cpu = os.path.join(ROOT, "oracle", "_ref", "libapd_cpu.so")
    assert os.path.exists(cpu), "run __graft_entry__.build()"
    lib = C.CDLL(cpu)
    assert hasattr(lib, "apd_cpu_run") and hasattr(lib, "apd_cpu_strong_pass")
    ref = os.path.join(ROOT, "oracle", "_ref", "libapd_ref.so")
    if os.path.exists(ref):
        # loading needs libcudart only, no device
        lib = C.CDLL(ref)
        for n in ("apdref_create", "apdref_upload", "apdref_run", "apdref_get", "apdref_get_outputs", "apdref_destroy"):
            assert hasattr(lib, n)
