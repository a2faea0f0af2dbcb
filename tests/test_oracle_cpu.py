"""CPU-only: the plain-C restatement (oracle/apd_cpu.c) against the golden fixtures, i.e. against dumps
of the reference's own CUDA build (tests/golden/make_golden.py). This is what pins the oracle.

Tolerances. The RNG stage is integer work: bit-exact. Everything downstream is fp32 computed on the GPU
with --use_fast_math (MUFU approximations, FTZ, hardware bilinear filter), which the C code follows with
IEEE operations and a model of the texture unit, so single costs agree to ~1e-4 and, because every output
is the result of discrete decisions on such costs, agreement is distributional: the
thresholds below are the measured agreement minus a safety margin.
"""
import numpy as np
import pytest

import golden_tools as G
from oracle import cpu_binding as CB

REL = 1e-4          # north_star tolerance on depth / normal / cost


def run_cpu(case, stage):
    p = G.oracle_params(case)
    st, _ = CB.run(case["images"], case["cameras"], p, depths=case["depths"], planes=case["planes"],
                   views=case["views"], states=case["states"], stage_end=stage)
    return st


def frac_close(a, b, tol=REL, relative=True):
    a, b = a.astype(np.float64), b.astype(np.float64)
    ok = np.isfinite(a) & np.isfinite(b)
    err = np.abs(a - b) / (np.maximum(np.abs(b), 1e-12) if relative else 1.0)
    same_nan = ~np.isfinite(a) & ~np.isfinite(b)
    return float(((err <= tol) & ok | same_nan).mean())


def test_rng_init_bit_exact():
    g, case = G.load("strong_first_64x48_s2")
    st = run_cpu(case, 0)
    assert np.array_equal(st.rng, g["s0_rng"])          # curand_init(seed, y, x) incl. the 2^67 subsequence skip


@pytest.mark.parametrize("name", ["strong_first_64x48_s2", "strong_geom_64x48_s3", "strong_refineinit_48x40_s2"])
def test_initial_planes_and_costs(name):
    g, case = G.load(name)
    st = run_cpu(case, 4)
    assert frac_close(st.planes[..., 3], g["s4_planes"][..., 3]) == 1.0
    assert np.abs(st.planes[..., :3] - g["s4_planes"][..., :3]).max() <= 1e-5
    # textureless rectangles (variance ~0.1) make single costs ill-conditioned there: ~4 % of pixels exceed 1e-3
    assert frac_close(st.costs, g["s4_costs"], tol=1e-3, relative=False) >= 0.93
    assert frac_close(st.costs, g["s4_costs"], tol=2e-2, relative=False) >= 0.985
    assert np.array_equal(st.states, g["s4_states"])
    assert (st.views == g["s4_views"]).mean() >= 0.99


def test_rng_stream_position_after_full_run():
    """Same number of draws per pixel as the reference: integer property, bit-exact."""
    for name, last in (("strong_first_64x48_s2", 14), ("strong_geom_64x48_s3", 14), ("strong_refineinit_48x40_s2", 19)):
        g, case = G.load(name)
        st = run_cpu(case, last)
        assert np.array_equal(st.rng, g[f"s{last}_rng"])


def test_one_iteration_distributional():
    g, case = G.load("strong_first_64x48_s2")
    st = run_cpu(case, 6)            # after K5, K6, K7
    assert frac_close(st.planes[..., 3], g["s6_planes"][..., 3], tol=1e-3) >= 0.90
    assert (st.views == g["s6_views"]).mean() >= 0.98
    assert (st.view_weights[..., :8] == g["s6_vw"]).all(-1).mean() >= 0.95


def test_refine_init_two_iterations():
    g, case = G.load("strong_refineinit_48x40_s2")
    st = run_cpu(case, 19)
    assert frac_close(st.planes[..., 3], g["planes"][..., 3], tol=1e-3) >= 0.99
    assert (st.states == g["states"]).mean() >= 0.97
    assert (st.views == g["views"]).mean() >= 0.99


def test_classification_with_geometric_term():
    g, case = G.load("strong_geom_64x48_s3")
    st = run_cpu(case, 14)
    assert (st.states == g["states"]).mean() >= 0.90
    assert set(np.unique(st.states)) <= {0, 1, 2}
    assert (st.states[:6] == 2).all() and (st.states[:, -6:] == 2).all()      # 6-px border -> UNKNOWN (APD.cu:2001)


# ---- adaptive patch deformation (WEAK pixels): K2-K4, K8-K10 --------------------------------------------------------
APD_FIXTURES = ["apd_geom_96x72_s3", "apd_init_96x72_s3_rot2"]


def run_cpu_apd(name, stage):
    g, case = G.load(name)
    p = G.oracle_params(case)
    st, extra, n = CB.run_apd(case["images"], case["cameras"], p, depths=case["depths"], planes=case["planes"],
                              views=case["views"], states=case["states"], seed=int(g["seed"]), stage_end=stage)
    return g, case, st, extra


@pytest.mark.parametrize("name", APD_FIXTURES)
def test_nearest_strong_point_bit_exact(name):
    g, case, st, ex = run_cpu_apd(name, 1)                 # K2 is integer work: exact
    weak = case["states"] == 0
    assert weak.sum() > 1000
    assert np.array_equal(ex["nearest"][weak], g["nearest"][weak])
    assert (ex["nearest"][~weak] == -1).all()


@pytest.mark.parametrize("name", APD_FIXTURES)
def test_anchor_search_and_reliability(name):
    g, case, st, ex = run_cpu_apd(name, 3)                 # K3 + K4
    weak = case["states"] == 0
    ref = g["anchors_compact"][g["anchors_map"][weak]]     # the reference stores anchors per WEAK index
    mine = ex["anchors"][weak]
    assert np.array_equal(mine[:, 0], ref[:, 0])           # slot 0 is the pixel itself
    same_set = np.array([set(map(tuple, a.tolist())) == set(map(tuple, b.tolist())) for a, b in zip(mine, ref)])
    # The curand draws are integer work and the search visits the same pixels; the ORDER of the three points that define
    # the RANSAC plane depends on their ~1e-7 residuals to their own plane, which fast-math decides differently.
    assert same_set.mean() >= 0.99
    assert (mine == ref).all(axis=(1, 2)).mean() >= 0.6
    assert (ex["reliable"][weak] == g["reliable"][weak]).mean() >= 0.995
    assert (st.states == g["s3_states"]).mean() >= 0.995   # NeigbourUpdate: unreliable WEAK -> UNKNOWN


@pytest.mark.parametrize("name,depth_ok", [("apd_geom_96x72_s3", 0.75), ("apd_init_96x72_s3_rot2", 0.95)])
def test_first_weak_update_distributional(name, depth_ok):
    g, case, st, ex = run_cpu_apd(name, 9)                 # ... K8 fit plane, K9 weak black of iteration 0
    assert (st.states == g["s9_states"]).mean() >= 0.995
    weak = g["s9_states"] == 0
    assert frac_close(st.planes[..., 3][weak], g["s9_planes"][..., 3][weak], tol=1e-3) >= depth_ok
    assert (st.views[weak] == g["s9_views"][weak]).mean() >= 0.98
    assert frac_close(st.costs[weak], g["s9_costs"][weak], tol=2e-2, relative=False) >= 0.90


@pytest.mark.parametrize("name", APD_FIXTURES)
def test_full_apd_pass_distributional(name):
    g, case, st, ex = run_cpu_apd(name, -1)
    assert (st.states == g["states"]).mean() >= 0.97
    assert frac_close(st.planes[..., 3], g["planes"][..., 3], tol=1e-2) >= 0.95
    assert frac_close(st.planes[..., 3], g["planes"][..., 3], tol=1e-3) >= 0.90
