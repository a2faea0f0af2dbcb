import numpy as np
from apd_mvs_b200.scene import make_scene, make_priors, make_cameras
from apd_mvs_b200 import shard


def test_cameras_are_consistent():
    cams = make_cameras(320, 240, 9)
    for c in cams:
        R = c["R"].reshape(3, 3).astype(np.float64)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)
        assert np.allclose(-R.T @ c["t"].astype(np.float64), c["c"], atol=1e-5)      # c = -R^T t (APD.cpp:75-77)
        assert c["width"] == 320 and c["height"] == 240


def test_scene_is_deterministic_and_has_exact_depth():
    a = make_scene(96, 64, 2)
    b = make_scene(96, 64, 2)
    assert np.array_equal(a["images"].numpy(), b["images"].numpy())
    d = a["depth"].numpy()
    assert d.min() > 3.0 and d.max() < 8.0
    assert 0.05 < a["weak_mask"].float().mean() < 0.7
    n = a["normal"].numpy()
    assert np.allclose(np.linalg.norm(n, axis=-1), 1.0, atol=1e-5) and (n[..., 2] < 0).all()


def test_priors_shapes_and_states():
    s = make_scene(96, 64, 3)
    p = make_priors(s)
    assert p["planes"].shape == (64, 96, 4) and p["depths"].shape == (4, 64, 96)
    assert set(np.unique(p["states"])) <= {0, 1, 2} and (p["states"][:6] == 2).all()
    assert (p["views"] == 0b111).all()


def test_view_ring_sharding():
    assert shard.view_order(0, 9, 12) == list(range(10))
    assert shard.view_order(5, 9, 12) == [5, 6, 7, 8, 9, 10, 11, 0, 1, 2]
    units = [u for r in range(4) for u in shard.units_of_rank(r, 4, 32)]
    assert sorted(units) == list(range(32)) and shard.units_of_rank(3, 4, 32)[:2] == [3, 7]
