"""CPU-only: the fusion oracle (the reference's unmodified APD.cpp behind oracle/shim_host) builds, loads and fuses a
synthetic scene with exact depth maps; the GPU fusion's C-ABI symbols are exported (tests/test_abi.py)."""
import numpy as np
import pytest

import fusion_tools as FT
from apd_mvs_b200.scene import make_scene

pytestmark = pytest.mark.skipif(not FT.ref_available(), reason="oracle/_ref/libapd_fusion_ref.so not built")


def test_reference_fusion_on_exact_depth_maps(tmp_path):
    V, W, H = 4, 96, 72
    sc = make_scene(W, H, V - 1, textureless=False)
    gray = sc["images"].numpy(); cams = sc["cameras"]; depth = sc["depth"].numpy()
    n0 = sc["normal"].numpy()
    normals = np.broadcast_to(n0[None], (V, H, W, 3)).copy()       # one plane dominates; elsewhere the angle test may reject
    weaks = np.ones((V, H, W), np.uint8)
    ids = [0, 1, 2, 3]
    FT.write_dense_folder(tmp_path, ids, FT.colour_images(gray), cams, depth, normals, weaks)
    problems = [(i, [j for j in ids if j != i]) for i in ids]
    xyz, bgr = FT.run_reference_fusion(tmp_path, problems)
    assert len(xyz) > 0.5 * W * H                                   # most of the first view is fused
    assert len(xyz) < V * W * H                                     # later views were masked by earlier ones
    # fused points lie on the scene (depth range of the generator) in front of camera 0
    assert np.isfinite(xyz).all() and (xyz[:, 2] > 2.5).all() and (xyz[:, 2] < 9.0).all()
    assert bgr.shape == (len(xyz), 3)
