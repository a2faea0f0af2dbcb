"""The reference's OWN main() (main.cpp, unmodified) run twice on the same dense_folder:
  oracle/_ref/apd_main_ref  = main.cpp + APD.cpp + APD.cu           (the reference program)
  oracle/_ref/apd_main_b200 = main.cpp + APD.cpp + facade + libapd_b200.so   (the integration of INTEGRATION.md)
with the same curand seed. Everything the program does - pair.txt parsing, image/camera loading, 4 passes per view with
results handed over through .dmb files, RunFusion, PLY export - must end in a byte-identical APD/APD.ply."""
import os
import shutil
import subprocess
import time

import numpy as np
import pytest

import fusion_tools as FT
from apd_mvs_b200.scene import make_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "apd_main_ref")
OUR_EXE = os.path.join(ROOT, "oracle", "_ref", "apd_main_b200")


def write_inputs(root, ids, bgr, cams, pairs):
    os.makedirs(root, exist_ok=True)
    V, H, W = bgr.shape[:3]
    zeros = np.zeros((V, H, W), np.float32)
    FT.write_dense_folder(root, ids, bgr, cams, zeros, np.zeros((V, H, W, 3), np.float32), np.zeros((V, H, W), np.uint8))
    shutil.rmtree(os.path.join(root, "APD"))               # main() creates it (main.cpp:146-147)
    with open(os.path.join(root, "pair.txt"), "w") as f:
        f.write(f"{len(pairs)}\n")
        for r, ss in pairs:
            f.write(f"{ids[r]}\n{len(ss)} " + " ".join(f"{ids[s]} {100.0 - k:.1f}" for k, s in enumerate(ss)) + "\n")


@pytest.mark.skipif(not (os.path.exists(REF_EXE) and os.path.exists(OUR_EXE)), reason="oracle/_ref/apd_main_* not built (make -C oracle main)")
def test_reference_main_with_and_without_the_facade(tmp_path):
    W, H, V = 320, 240, 4
    sc = make_scene(W, H, V - 1, device="cuda")
    bgr = FT.colour_images(sc["images"].cpu().numpy())
    ids = [0, 1, 2, 3]
    pairs = [(r, [(r + k) % V for k in (1, 2, 3)]) for r in range(V)]
    a, b = str(tmp_path / "ref"), str(tmp_path / "ours")
    write_inputs(a, ids, bgr, sc["cameras"], pairs); write_inputs(b, ids, bgr, sc["cameras"], pairs)
    t0 = time.perf_counter()
    keep = dict(os.environ, APD_SHIM_KEEP_FILES="1")          # main() would delete the intermediate matrix files (main.cpp:219-226)
    r1 = subprocess.run([REF_EXE, a, "0"], capture_output=True, text=True, timeout=600, env=keep)
    t1 = time.perf_counter()
    r2 = subprocess.run([OUR_EXE, b, "0"], capture_output=True, text=True, timeout=600, env=dict(keep, APD_SEED="1234567"))
    t2 = time.perf_counter()
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert "All done" in r1.stdout and "All done" in r2.stdout
    pa, pb = open(os.path.join(a, "APD", "APD.ply"), "rb").read(), open(os.path.join(b, "APD", "APD.ply"), "rb").read()
    xyz, _ = FT.read_ply(os.path.join(a, "APD", "APD.ply"))
    assert len(xyz) > 0.2 * W * H * V / 2
    assert pa == pb, "APD.ply differs between the reference program and the facade build"
    # the matrix files the REFERENCE PROGRAM wrote (WriteBinMat, APD.cpp:132-176: depths.dmb CV_32FC1, normals.dmb CV_32FC3) are read by
    # the product's reader (apd_io_read_mat) and equal what the facade build wrote, as are weak.bin (CV_8UC1) and selected_views.bin (CV_32SC1)
    from apd_mvs_b200 import io as IO
    import glob
    folders = sorted(glob.glob(os.path.join(a, "APD", "*", "depths.dmb")))
    assert len(folders) == V
    for fa in folders:
        fb = fa.replace(a, b, 1)
        d_ref, d_our = IO.ReadBinMat(fa), IO.ReadBinMat(fb)
        assert d_ref.shape == (H, W) and d_ref.dtype == np.float32 and np.isfinite(d_ref).all() and (d_ref > 0).mean() > 0.5
        assert np.array_equal(d_ref.view(np.uint32), d_our.view(np.uint32))
        n_ref = IO.ReadBinMat(fa.replace("depths.dmb", "normals.dmb"))
        assert n_ref.shape == (H, W, 3) and n_ref.dtype == np.float32
        assert np.array_equal(n_ref.view(np.uint32), IO.ReadBinMat(fb.replace("depths.dmb", "normals.dmb")).view(np.uint32))
        w_ref = IO.ReadBinMat(fa.replace("depths.dmb", "weak.bin")); v_ref = IO.ReadBinMat(fa.replace("depths.dmb", "selected_views.bin"))
        assert w_ref.shape == (H, W) and w_ref.dtype == np.uint8 and v_ref.shape == (H, W) and v_ref.dtype.itemsize == 4
        assert np.array_equal(w_ref, IO.ReadBinMat(fb.replace("depths.dmb", "weak.bin")))
        assert np.array_equal(v_ref, IO.ReadBinMat(fb.replace("depths.dmb", "selected_views.bin")))
    print(f"reference program {t1 - t0:.2f} s, facade build {t2 - t1:.2f} s, {len(xyz)} fused points")
