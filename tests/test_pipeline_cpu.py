"""CPU tests of the pass-scheduler oracle (oracle/ref_pipeline.py): the restated host logic of main.cpp /
APD::InuputInitialization against direct readings of the reference and against the installed OpenCV."""
import math

import numpy as np
import pytest

from oracle import ref_pipeline as RP


def test_round_num_matches_main_cpp():
    # main.cpp:72-88 and the sizes quoted in SURVEY §3.1
    assert RP.compute_round_num(6221, 4146) == 4
    assert RP.compute_round_num(3111, 2074) == 3
    assert RP.compute_round_num(1920, 1080) == 2
    assert RP.compute_round_num(1000, 700) == 1
    assert RP.compute_round_num(1001, 700) == 2
    assert [RP.scale_size(4, i) for i in range(4)] == [8, 4, 2, 1]


def test_scaled_size_rounds_half_away():
    assert RP.scaled_size(6221, 4146, 2) == (3111, 2073)     # 3110.5 -> 3111
    assert RP.scaled_size(3111, 2074, 2) == (1556, 1037)
    assert RP.scaled_size(1041, 781, 2) == (521, 391)
    assert RP.scaled_size(1041, 781, 1) == (1041, 781)
    assert RP.scaled_size(6221, 4146, 8) == (778, 518)


class _P:
    pass


def test_pass_params_schedule():
    mk = lambda: _P()
    p = RP.pass_params(mk, 0, 0)
    assert (p.state, p.use_APD, p.geom_consistency, p.weak_peak_radius, p.max_iterations) == (RP.FIRST_INIT, 0, 0, 6, 3)
    p = RP.pass_params(mk, 0, 2)
    assert (p.state, p.use_APD, p.geom_consistency, p.weak_peak_radius) == (RP.REFINE_ITER, 0, 1, 2)
    p = RP.pass_params(mk, 2, 0)
    assert (p.state, p.use_APD, p.geom_consistency, p.rotate_time) == (RP.REFINE_INIT, 1, 0, 4)
    assert p.ransac_threshold == float(np.float32(0.01 - 2 * 0.00125))
    p = RP.pass_params(mk, 1, 1)
    assert (p.state, p.weak_peak_radius, p.rotate_time) == (RP.REFINE_ITER, 4, 2)
    assert [RP.pass_params(mk, 1, j).weak_peak_radius for j in (1, 2, 3)] == [4, 2, 2]


def _rescale_loop(src, dw, dh):
    """APD.cpp:752-774 written as the double loop it is."""
    sh, sw = src.shape[:2]
    scale_x = np.float32(dw) / np.float32(sw)
    scale_y = np.float32(dh) / np.float32(sh)
    out = np.zeros((dh, dw) + src.shape[2:], src.dtype)
    for r in range(dh):
        o_r = int(np.float32(r) / scale_x)
        for c in range(dw):
            o_c = int(np.float32(c) / scale_y)
            if o_r < 0 or o_c < 0 or o_r >= sh or o_c >= sw:
                continue
            out[r, c] = src[o_r, o_c]
    return out


@pytest.mark.parametrize("sw,sh,dw,dh", [(52, 39, 104, 78), (53, 40, 105, 79), (40, 53, 79, 105), (60, 20, 119, 41)])
def test_rescale_nearest_matches_reference_loop(sw, sh, dw, dh):
    rng = np.random.default_rng(3)
    src = rng.random((sh, sw)).astype(np.float32)
    assert np.array_equal(RP.rescale_nearest(src, dw, dh), _rescale_loop(src, dw, dh))
    src3 = rng.random((sh, sw, 3)).astype(np.float32)
    assert np.array_equal(RP.rescale_nearest(src3, dw, dh), _rescale_loop(src3, dw, dh))
    srcu = rng.integers(0, 3, (sh, sw)).astype(np.uint8)
    assert np.array_equal(RP.rescale_nearest(srcu, dw, dh), _rescale_loop(srcu, dw, dh))


def test_rescale_identity_returns_input():
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    assert RP.rescale_nearest(a, 4, 3) is a


@pytest.mark.parametrize("w,h,scale", [(1041, 781, 2), (1040, 780, 2), (777, 555, 2), (1111, 999, 4), (1040, 780, 4)])
def test_resize_restatement_against_installed_opencv(w, h, scale):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(w)
    yy, xx = np.mgrid[0:h, 0:w]
    img = (127.5 + 100 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + rng.uniform(-20, 20, (h, w))).astype(np.float32)
    dw, dh = RP.scaled_size(w, h, scale)
    mine = RP.resize_linear(img, dw, dh)
    ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    assert mine.shape == ref.shape
    # OpenCV's optimised builds (IPP / FMA) differ from its generic path in the last bits: 1.5e-4 of full scale
    assert np.abs(mine - ref).max() <= 1.5e-4 * 255.0


def test_resize_half_is_box_average():
    img = np.arange(48, dtype=np.float32).reshape(6, 8)
    out = RP.resize_linear(img, 4, 3)
    assert np.array_equal(out, (img[0::2, 0::2] + img[0::2, 1::2] + img[1::2, 0::2] + img[1::2, 1::2]) * np.float32(0.25))
