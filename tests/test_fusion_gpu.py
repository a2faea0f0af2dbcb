"""GPU parity of the depth-map fusion (include/apd_fusion.h) against the reference's UNMODIFIED RunFusion
(APD.cpp:826-977 compiled behind oracle/shim_host, run on the host CPU) on the same depth / normal / state maps,
which come from a real run of the pass schedule. Bar: the same points in the same order; coordinates and colours
bit-identical, pinned at ZERO differing points. (The only operations that could legitimately differ between glibc and CUDA are
acosf / expf in their last bit; on these inputs no value sits on a threshold.)"""
import numpy as np
import pytest

import fusion_tools as FT
from apd_mvs_b200 import engine as E
from apd_mvs_b200 import fusion as F
from apd_mvs_b200 import pipeline as P
from apd_mvs_b200.scene import make_scene

pytestmark = pytest.mark.gpu


def depth_maps(W, H, V, n_src):
    sc = make_scene(W, H, V - 1, device="cuda")
    images, cams = sc["images"].cpu().numpy(), sc["cameras"]
    pairs = P.ring_pairs(V, n_src)
    scene = P.Scene(images, cams, pairs, seed=11)
    scene.Run()
    out = ([scene.Depth(v) for v in range(V)], [scene.Normal(v) for v in range(V)], [scene.States(v) for v in range(V)])
    scene.close()
    return images, cams, pairs, out


def fuse_gpu(bgr, cams, pairs, depths, normals, states, blocks=None, variant="eth"):
    V, H, W = len(depths), depths[0].shape[0], depths[0].shape[1]
    fu = F.Fusion(V, W, H)
    for v in range(V):
        fu.SetView(v, bgr[v], cams[v], depths[v], normals[v], states[v], None if blocks is None else blocks[v])
    for r, s in pairs:
        fu.AddProblem(r, s)
    xyz, col = fu.RunFusion(variant)
    return fu, xyz, col


@pytest.mark.skipif(not FT.ref_available(), reason="oracle/_ref/libapd_fusion_ref.so not built")
@pytest.mark.parametrize("W,H,V,n_src", [(320, 240, 4, 3), (401, 301, 5, 2)])
def test_fusion_matches_reference(tmp_path, W, H, V, n_src):
    images, cams, pairs, (depths, normals, states) = depth_maps(W, H, V, n_src)
    bgr = FT.colour_images(images)
    ids = [7 + 3 * v for v in range(V)]
    FT.write_dense_folder(tmp_path, ids, bgr, cams, depths, normals, states)
    ref_xyz, ref_bgr = FT.run_reference_fusion(tmp_path, [(ids[r], [ids[s] for s in ss]) for r, ss in pairs])
    fu, xyz, col = fuse_gpu(bgr, cams, pairs, depths, normals, states)
    assert len(ref_xyz) > 0.3 * W * H
    # pinned at zero differing points: same points, same order, same coordinates and colours (the fusion TU is compiled with
    # IEEE arithmetic; on these inputs no acosf / expf value sits on a threshold, so glibc and CUDA agree everywhere)
    assert len(xyz) == len(ref_xyz), f"{len(xyz)} vs {len(ref_xyz)} points"
    assert np.array_equal(xyz.view(np.uint32), ref_xyz.view(np.uint32))
    assert np.array_equal(col.astype(np.uint8), ref_bgr)          # ExportPointCloud truncates to uchar
    # the PLY writer produces the reference's file layout
    ply = tmp_path / "ours.ply"
    fu.ExportPointCloud(ply)
    pxyz, pbgr = FT.read_ply(ply)
    assert np.array_equal(pxyz.view(np.uint32), xyz.view(np.uint32)) and np.array_equal(pbgr, col.astype(np.uint8))
    t = fu.Timing()
    assert t["gpu_ms"] > 0 and t["max_rounds"] >= 1
    fu.close()


@pytest.mark.skipif(not FT.ref_available(), reason="oracle/_ref/libapd_fusion_ref.so not built")
@pytest.mark.parametrize("variant,code", [("tat_intermediate", 1), ("tat_advanced", 2)])
def test_tanks_and_temples_variants_match_reference(tmp_path, variant, code):
    """RunFusion_TAT_Intermediate / RunFusion_TAT_advanced (APD.cpp:979-1296), including the carry-over of a source's last
    measurement to later pixels (`diff` is declared per view, not per pixel): same points, same order, same bits."""
    W, H, V, n_src = 352, 264, 5, 4
    images, cams, pairs, (depths, normals, states) = depth_maps(W, H, V, n_src)
    bgr = FT.colour_images(images)
    ids = [3 + 2 * v for v in range(V)]
    FT.write_dense_folder(tmp_path, ids, bgr, cams, depths, normals, states)
    ref_xyz, ref_bgr = FT.run_reference_fusion(tmp_path, [(ids[r], [ids[s] for s in ss]) for r, ss in pairs], variant=code)
    fu, xyz, col = fuse_gpu(bgr, cams, pairs, depths, normals, states, variant=variant)
    assert len(ref_xyz) > 0.05 * W * H, len(ref_xyz)
    assert len(xyz) == len(ref_xyz), f"{len(xyz)} vs {len(ref_xyz)} points"
    assert np.array_equal(xyz.view(np.uint32), ref_xyz.view(np.uint32))
    assert np.array_equal(col.astype(np.uint8), ref_bgr)
    fu.close()


def test_fusion_is_deterministic_and_order_preserving():
    images, cams, pairs, (depths, normals, states) = depth_maps(256, 192, 3, 2)
    bgr = FT.colour_images(images)
    fu1, xyz1, col1 = fuse_gpu(bgr, cams, pairs, depths, normals, states)
    fu2, xyz2, col2 = fuse_gpu(bgr, cams, pairs, depths, normals, states)
    assert np.array_equal(xyz1.view(np.uint32), xyz2.view(np.uint32)) and np.array_equal(col1, col2)
    # block masks < 128 remove reference pixels (APD.cpp:905-907)
    blocks = [np.full((192, 256), 255, np.uint8) for _ in range(3)]
    blocks[0][:, :128] = 0
    fu3, xyz3, _ = fuse_gpu(bgr, cams, pairs, depths, normals, states, blocks)
    assert 0 < len(xyz3) < len(xyz1)
    for f in (fu1, fu2, fu3):
        f.close()


def test_fusion_argument_errors():
    fu = F.Fusion(3, 64, 48)
    with pytest.raises(E.ApdError):
        fu.AddProblem(0, [0])                  # a view cannot be its own source
    with pytest.raises(E.ApdError):
        fu.AddProblem(0, [5])
    fu.AddProblem(0, [1, 2])
    with pytest.raises(E.ApdError):
        fu.RunFusion()                         # views not set
    fu.close()
