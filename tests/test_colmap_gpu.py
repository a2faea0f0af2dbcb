"""The converter's view-selection scoring on the GPU (torch tensors on cuda:0) gives the reference converter's pair.txt."""
import os
import types

import pytest

from apd_mvs_b200 import colmap2mvsnet as C2M

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_pair_list_from_gpu_scores(tmp_path):
    a = types.SimpleNamespace(dense_folder=os.path.join(HERE, "golden", "colmap_scene"), save_folder=str(tmp_path), max_d=192, interval_scale=1,
                              scale_factor=1, theta0=5, sigma1=1, sigma2=10, model_ext=".bin", device="cuda:0")
    C2M.processing_single_scene(a)
    want = os.path.join(HERE, "golden", "colmap_expected_bin")
    assert open(tmp_path / "pair.txt").read() == open(os.path.join(want, "pair.txt")).read()
    assert open(tmp_path / "cams" / "00000004_cam.txt").read() == open(os.path.join(want, "cams", "00000004_cam.txt")).read()
