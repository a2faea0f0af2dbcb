"""GPU parity of the scene layer (include/apd_scene.h: pass scheduler + persistent device cache) against the
reference's driver loop restated around the unmodified reference PatchMatch (oracle/ref_pipeline.py +
oracle/_ref/libapd_ref.so). Single-GPU, pair-list order, so every (problem, pass) sees exactly the inputs it sees
in the reference and the whole 4*round_num-pass pipeline must come out bit-identical."""
import ctypes as C

import numpy as np
import pytest

import parity_tools as T
from apd_mvs_b200 import engine as E
from apd_mvs_b200 import pipeline as P
from apd_mvs_b200.scene import make_scene
from oracle import ref_pipeline as RP

pytestmark = pytest.mark.gpu


def ref_available():
    from oracle import ref_binding
    return ref_binding.available()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def run_reference_patchmatch(images, cams, params, depths, planes, views, states, seed):
    from oracle.ref_binding import RefAPD
    ref = RefAPD(images, cams, T.clone_params(params), depths=depths, planes=planes, views=views, states=states, seed=seed)
    ref.run()
    out = ref.outputs()
    ref.close()
    return out


def make_views(W, H, n_views):
    sc = make_scene(W, H, n_views - 1, device="cuda")
    return sc["images"].cpu().numpy(), sc["cameras"]


@pytest.mark.parametrize("W,H", [(1041, 781), (1040, 780)])
def test_scaled_images_bit_exact(W, H):
    images, cams = make_views(W, H, 3)
    sc = P.Scene(images, cams, P.ring_pairs(3, 2))
    assert sc.ComputeRoundNum() == RP.compute_round_num(W, H) == 2
    assert sc.RoundSize(0) == RP.scaled_size(W, H, 2) and sc.RoundSize(1) == (W, H)
    for v in range(3):
        w, h = sc.RoundSize(0)
        assert np.array_equal(bits(sc.ScaledImage(0, v)), bits(RP.resize_linear(images[v], w, h)))
        assert np.array_equal(bits(sc.ScaledImage(1, v)), bits(images[v]))
    sc.close()


def test_pass_params_match_schedule():
    images, cams = make_views(64, 48, 2)
    big = np.zeros((2, 48, 4200), np.float32)       # 4200 px wide -> 4 rounds; only the schedule is queried
    cams_big = cams.copy(); cams_big["width"] = 4200
    sc = P.Scene(big, cams_big, [(0, [1])])
    assert sc.ComputeRoundNum() == 4
    for i in range(4):
        for ps in range(4):
            a, b = sc.PassParams(i, ps), RP.pass_params(E.default_params, i, ps)
            for f in ("max_iterations", "top_k", "geom_consistency", "use_APD", "weak_peak_radius", "rotate_time", "state"):
                assert getattr(a, f) == getattr(b, f), (i, ps, f)
            assert np.float32(a.ransac_threshold) == np.float32(b.ransac_threshold)
    sc.close()


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref/libapd_ref.so not built")
@pytest.mark.parametrize("W,H,n_views,n_src", [(1041, 781, 3, 2), (1040, 600, 4, 3)])
def test_whole_pipeline_bit_exact(W, H, n_views, n_src):
    images, cams = make_views(W, H, n_views)
    pairs = P.ring_pairs(n_views, n_src)
    ref = RP.RefPipeline(images, cams, pairs, E.default_params, run_reference_patchmatch, seed=4242)
    sc = P.Scene(images, cams, pairs, seed=4242)
    assert sc.ComputeRoundNum() == ref.rounds == 2
    for i in range(ref.rounds):
        for ps in range(4):
            ref.run_pass(i, ps)
            sc.RunPass(i, ps)
            for v in range(n_views):
                r = ref.results[v]
                assert sc.ResultSize(v) == (r["depth"].shape[1], r["depth"].shape[0])
                assert np.array_equal(bits(sc.Depth(v)), bits(r["depth"])), (i, ps, v, "depth")
                assert np.array_equal(bits(sc.Normal(v)), bits(r["normal"])), (i, ps, v, "normal")
                assert np.array_equal(sc.States(v), r["weak"]), (i, ps, v, "states")
                assert np.array_equal(sc.SelectedViews(v), r["views"]), (i, ps, v, "views")
    t = sc.Timing()
    assert t["launches"] > 0 and t["patchmatch_ms"] > 0
    # the final maps are usable: most pixels carry an in-range depth
    d = sc.Depth(0)
    assert (d > 0).mean() > 0.8
    sc.close()


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref/libapd_ref.so not built")
def test_unequal_source_counts_smaller_first():
    """pair.txt drops sources with score <= 0 (main.cpp:42-44), so problems of one round have different numbers of
    images. The depth-map array of the round's engine is created lazily by the first geometric problem: it must hold the
    round's maximum, not that problem's count (apd_engine.cu make_layered)."""
    images, cams = make_views(600, 448, 4)
    pairs = [(0, [1]), (1, [2, 3, 0]), (2, [3, 0]), (3, [0, 1, 2])]
    ref = RP.RefPipeline(images, cams, pairs, E.default_params, run_reference_patchmatch, seed=99)
    sc = P.Scene(images, cams, pairs, seed=99)
    for ps in range(4):
        ref.run_pass(0, ps)
        sc.RunPass(0, ps)
    for v in range(4):
        r = ref.results[v]
        assert np.array_equal(bits(sc.Depth(v)), bits(r["depth"])), v
        assert np.array_equal(sc.States(v), r["weak"]) and np.array_equal(sc.SelectedViews(v), r["views"]), v
    sc.close()


def test_run_equals_pass_by_pass():
    images, cams = make_views(1000, 512, 3)            # one round (max size <= 1000)
    pairs = P.ring_pairs(3, 2)
    a = P.Scene(images, cams, pairs, seed=7)
    b = P.Scene(images, cams, pairs, seed=7)
    a.Run()
    for ps in range(4):
        b.RunPass(0, ps)
    for v in range(3):
        assert np.array_equal(bits(a.Depth(v)), bits(b.Depth(v)))
        assert np.array_equal(a.States(v), b.States(v))
    a.close(); b.close()


def test_pass_order_is_enforced():
    images, cams = make_views(256, 192, 3)
    sc = P.Scene(images, cams, P.ring_pairs(3, 2))
    with pytest.raises(E.ApdError):
        sc.RunPass(0, 1)                                  # geometric pass before any depth map exists
    with pytest.raises(E.ApdError):
        sc.Depth(0)
    sc.RunPass(0, 0)
    assert sc.Depth(0).shape == (192, 256)
    sc.close()


def test_bad_arguments():
    images, cams = make_views(64, 48, 2)
    with pytest.raises(E.ApdError):
        P.Scene(images, cams, [(0, [5])])                 # source view out of range
    L = P._bind(E.lib())
    h = C.c_void_p(None)
    assert L.apd_scene_create(C.byref(h), 0, 1, 64, 48, 0) == -4   # APD_E_LIMIT: fewer than 2 views
    assert L.apd_scene_run(None) == -1


def test_dense_folder_command_line(tmp_path):
    """main.cpp's command line on a synthetic dense_folder: pair.txt + JPEGs + _cam.txt in, the four result files out."""
    cv2 = pytest.importorskip("cv2")
    from apd_mvs_b200 import io as IO, main as M
    W, H, V = 320, 240, 3
    images, cams = make_views(W, H, V)
    ids = [3, 10, 42]                                       # image ids need not be dense
    (tmp_path / "images").mkdir(); (tmp_path / "cams").mkdir()
    for k, image_id in enumerate(ids):
        name = IO.ToFormatIndex(image_id)
        cv2.imwrite(str(tmp_path / "images" / (name + ".jpg")), np.clip(images[k], 0, 255).astype(np.uint8))
        c = cams[k]
        R, t, K = c["R"].reshape(3, 3), c["t"], c["K"].reshape(3, 3)
        txt = "extrinsic\n" + "".join(" ".join(repr(float(x)) for x in list(R[i]) + [t[i]]) + "\n" for i in range(3)) + "0.0 0.0 0.0 1.0\n\n"
        txt += "intrinsic\n" + "".join(" ".join(repr(float(x)) for x in K[i]) + "\n" for i in range(3))
        txt += f"\n{float(c['depth_min'])!r} 0.01 192 {float(c['depth_max'])!r}\n"
        (tmp_path / "cams" / (name + "_cam.txt")).write_text(txt)
    (tmp_path / "pair.txt").write_text("3\n3\n2 10 9.0 42 8.0\n10\n3 42 7.0 3 6.0 7 0.0\n42\n2 3 5.0 10 4.0\n")
    assert M.main([str(tmp_path)]) == 0
    got_ids, imgs, cams2, pairs = M.load_dense_folder(str(tmp_path))
    assert got_ids == ids and pairs == [(0, [1, 2]), (1, [2, 0]), (2, [0, 1])]
    assert np.array_equal(cams2["R"], cams["R"]) and np.array_equal(cams2["K"], cams["K"]) and np.array_equal(cams2["t"], cams["t"])
    sc = P.Scene(imgs, cams2, pairs)
    sc.Run()
    for k, image_id in enumerate(ids):
        out = tmp_path / "APD" / IO.ToFormatIndex(image_id)
        d = IO.ReadBinMat(out / "depths.dmb"); n = IO.ReadBinMat(out / "normals.dmb")
        w = IO.ReadBinMat(out / "weak.bin"); v = IO.ReadBinMat(out / "selected_views.bin")
        assert d.shape == (H, W) and n.shape == (H, W, 3) and w.dtype == np.uint8 and v.shape == (H, W)
        assert np.array_equal(bits(d), bits(sc.Depth(k))) and np.array_equal(bits(n), bits(sc.Normal(k)))
        assert np.array_equal(w, sc.States(k)) and np.array_equal(v.view(np.uint32), sc.SelectedViews(k))
        assert (d > 0).mean() > 0.7
    sc.close()
    import fusion_tools as FT
    xyz, bgr = FT.read_ply(tmp_path / "APD" / "APD.ply")           # RunFusion output (main.cpp:219)
    assert len(xyz) > 0.15 * W * H and np.isfinite(xyz).all()


def test_schedule_and_fusion_against_committed_golden():
    """tests/golden/pipeline_1010x64_v3.json (made from the reference by tests/golden/make_golden_pipeline.py): the whole
    2-round schedule and the fusion of its maps, compared by SHA-256 - needs no oracle library at run time."""
    import hashlib
    import json
    import os
    import fusion_tools as FT
    from apd_mvs_b200 import fusion as F
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pipeline_1010x64_v3.json")))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    from apd_mvs_b200.scene import CAMERA_DTYPE
    inp = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pipeline_1010x64_v3_inputs.npz"))
    images = inp["images"].astype(np.float32)
    cams = np.frombuffer(inp["cameras"].tobytes(), dtype=CAMERA_DTYPE).copy()
    assert sha(images) == g["images_sha"]
    pairs = P.ring_pairs(g["views"], g["src"])
    sc = P.Scene(images, cams, pairs, seed=g["seed"])
    sc.Run()
    for v, r in enumerate(g["results"]):
        assert sha(sc.Depth(v)) == r["depth"], (v, "depth")
        assert sha(sc.Normal(v)) == r["normal"], (v, "normal")
        assert sha(sc.States(v)) == r["states"] and sha(sc.SelectedViews(v)) == r["views"], v
        assert np.allclose(sc.Depth(v)[::16, ::101].ravel()[:12], r["depth_samples"], rtol=0, atol=0)
    fu = F.Fusion(g["views"], g["W"], g["H"])
    bgr = FT.colour_images(images)
    for v in range(g["views"]):
        fu.SetView(v, bgr[v], cams[v], sc.Depth(v), sc.Normal(v), sc.States(v))
    for r, s in pairs:
        fu.AddProblem(r, s)
    xyz, col = fu.RunFusion()
    assert len(xyz) == g["fusion"]["points"]
    assert sha(xyz) == g["fusion"]["xyz"] and sha(col.astype(np.uint8)) == g["fusion"]["bgr"]
    fu.close(); sc.close()


def test_checkpoint_and_resume(tmp_path):
    """Stop after round 0, write the four files of every view, start a NEW scene from them and finish: identical to the
    uninterrupted schedule (the files are the reference's hand-over between passes)."""
    from apd_mvs_b200 import io as IO
    images, cams = make_views(1040, 300, 3)
    pairs = P.ring_pairs(3, 2)
    whole = P.Scene(images, cams, pairs, seed=21)
    whole.Run()
    first = P.Scene(images, cams, pairs, seed=21)
    for ps in range(4):
        first.RunPass(0, ps)
    for v in range(3):
        IO.WriteBinMat(tmp_path / f"d{v}.dmb", first.Depth(v)); IO.WriteBinMat(tmp_path / f"n{v}.dmb", first.Normal(v))
        IO.WriteBinMat(tmp_path / f"w{v}.bin", first.States(v)); IO.WriteBinMat(tmp_path / f"s{v}.bin", first.SelectedViews(v))
    first.close()
    second = P.Scene(images, cams, pairs, seed=21)
    for v in range(3):
        second.SetResult(v, IO.ReadBinMat(tmp_path / f"d{v}.dmb"), IO.ReadBinMat(tmp_path / f"n{v}.dmb"),
                         IO.ReadBinMat(tmp_path / f"w{v}.bin"), IO.ReadBinMat(tmp_path / f"s{v}.bin"))
    for ps in range(4):
        second.RunPass(1, ps)
    for v in range(3):
        assert np.array_equal(bits(second.Depth(v)), bits(whole.Depth(v)))
        assert np.array_equal(bits(second.Normal(v)), bits(whole.Normal(v)))
        assert np.array_equal(second.States(v), whole.States(v)) and np.array_equal(second.SelectedViews(v), whole.SelectedViews(v))
    whole.close(); second.close()
