"""TEST INFRASTRUCTURE (oracle): the reference's driver loop restated in numpy around the UNMODIFIED reference
PatchMatch (oracle/_ref/libapd_ref.so). Used by tests/ and tools/ as the checker of the scene layer
(include/apd_scene.h); never imported by the product.

Follows, line by line:
  main.cpp:72-88    ComputeRoundNum
  main.cpp:168-217  round / pass schedule and per-pass parameters
  main.cpp:91-138   ProcessProblem (depth range test, UNKNOWN marking, the four result files)
  APD.cpp:399-583   InuputInitialization (image + camera scaling, depth maps, weak info, prior planes, selected views)
  APD.cpp:752-774   RescaleMatToTargetSize (nearest; row index divided by the x scale, column by the y scale)
cv::resize(INTER_LINEAR, CV_32FC1) is restated from OpenCV's generic C++ path (imgproc/resize.cpp); OpenCV itself
is a third-party dependency without a pinned version (CMakeLists.txt:9) and its optimised builds (IPP, AVX2 FMA)
differ from that path in the last bits: tests/test_pipeline_cpu.py pins this restatement against the installed
cv2 to 1.5e-4 relative, and the scene layer against this restatement bit-exactly. Result files are kept as arrays.
"""
from __future__ import annotations

import math
import numpy as np

FIRST_INIT, REFINE_INIT, REFINE_ITER = 0, 1, 2
WEAK, STRONG, UNKNOWN = 0, 1, 2
f32 = np.float32


def compute_round_num(W: int, H: int, limit: int = 1000) -> int:
    max_size, rounds = max(W, H), 1
    while max_size > limit:
        max_size //= 2
        rounds += 1
    return rounds


def scale_size(rounds: int, i: int) -> int:
    return int(math.pow(2, rounds - 1 - i))


def scaled_size(W: int, H: int, scale: int):
    if scale == 1:
        return W, H
    factor = f32(1.0) / f32(scale)

    def rnd(v):  # std::round of a float: half away from zero
        v = float(f32(v) * factor)
        return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))
    return rnd(W), rnd(H)


def pass_params(make_params, i: int, pass_: int):
    """main.cpp:171-211. make_params() returns a default PatchMatchParams-like object."""
    p = make_params()
    p.max_iterations = 3
    if i == 0:
        p.use_APD = 0
    else:
        p.use_APD = 1
        p.ransac_threshold = float(f32(0.01 - i * 0.00125))
        p.rotate_time = min(int(math.pow(2, i)), 4)
    if pass_ == 0:
        p.state = FIRST_INIT if i == 0 else REFINE_INIT
        p.geom_consistency = 0
        p.weak_peak_radius = 6
    else:
        j = pass_ - 1
        p.state = REFINE_ITER
        p.geom_consistency = 1
        p.weak_peak_radius = max(4 - 2 * j, 2)
    return p


def _table(dn: int, sn: int, horizontal: bool):
    inv_scale = float(dn) / float(sn)
    sc = 1.0 / inv_scale
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * sc - 0.5).astype(f32)
    o = np.floor(f).astype(np.int64)
    f = (f - o.astype(f32)).astype(f32)
    if horizontal:
        lo = o < 0; f[lo] = 0; o[lo] = 0
        hi = o >= sn - 1; f[hi] = 0; o[hi] = sn - 1
    return o, f


def resize_linear(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_LINEAR) for CV_32FC1, generic path."""
    img = np.ascontiguousarray(img, dtype=f32)
    sh, sw = img.shape
    if sw == 2 * dw and sh == 2 * dh:      # both scales exactly 2: cv::resize takes the 2x2 box average
        a, b = img[0::2, 0::2], img[0::2, 1::2]
        c, d = img[1::2, 0::2], img[1::2, 1::2]
        return (((a + b).astype(f32) + (c + d).astype(f32)).astype(f32) * f32(0.25)).astype(f32)
    xo, xw = _table(dw, sw, True)
    yo, yw = _table(dh, sh, False)
    x1 = np.minimum(xo + 1, sw - 1)
    a1 = xw[None, :]; a0 = (f32(1.0) - xw)[None, :].astype(f32)
    y0 = np.clip(yo, 0, sh - 1); y1 = np.clip(yo + 1, 0, sh - 1)
    b1 = yw[:, None]; b0 = (f32(1.0) - yw)[:, None].astype(f32)

    def hrow(rows):
        return ((rows[:, xo] * a0).astype(f32) + (rows[:, x1] * a1).astype(f32)).astype(f32)
    h0, h1 = hrow(img[y0]), hrow(img[y1])
    return ((h0 * b0).astype(f32) + (h1 * b1).astype(f32)).astype(f32)


def rescale_nearest(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """RescaleMatToTargetSize<T>, APD.cpp:752-774 (unwritten elements -> 0)."""
    sh, sw = src.shape[:2]
    if sw == dw and sh == dh:
        return src
    scale_x = f32(dw) / f32(sw)
    scale_y = f32(dh) / f32(sh)
    o_r = (np.arange(dh, dtype=f32) / scale_x).astype(np.int64)      # sic: rows by scale_x
    o_c = (np.arange(dw, dtype=f32) / scale_y).astype(np.int64)      # sic: columns by scale_y
    ok_r, ok_c = (o_r >= 0) & (o_r < sh), (o_c >= 0) & (o_c < sw)
    out = np.zeros((dh, dw) + src.shape[2:], dtype=src.dtype)
    rr, cc = np.nonzero(ok_r)[0], np.nonzero(ok_c)[0]
    out[np.ix_(rr, cc)] = src[np.ix_(o_r[rr], o_c[cc])]
    return out


class RefPipeline:
    """main() of the reference on in-memory views. `run_patchmatch(images, cams, params, depths, planes, views,
    states, seed) -> (planes[H,W,4], states[H,W], views[H,W])` executes APD::RunPatchMatch (oracle/_ref)."""

    def __init__(self, images, cameras, pairs, make_params, run_patchmatch, seed: int = 1234567, round_limit: int = 1000):
        self.images = np.ascontiguousarray(images, dtype=f32)
        self.n_views, self.H, self.W = self.images.shape
        self.cameras = cameras.copy()
        self.pairs = pairs
        self.make_params = make_params
        self.run_patchmatch = run_patchmatch
        self.seed = seed
        self.rounds = compute_round_num(self.W, self.H, round_limit)
        self.results = [None] * self.n_views       # dict(depth, normal, weak, views) = the four files of a view
        self._scaled_cache = {}

    def scaled_image(self, i: int, view: int):
        key = (i, view)
        if key not in self._scaled_cache:
            w, h = scaled_size(self.W, self.H, scale_size(self.rounds, i))
            self._scaled_cache[key] = self.images[view] if (w, h) == (self.W, self.H) else resize_linear(self.images[view], w, h)
        return self._scaled_cache[key]

    def process_problem(self, i: int, pass_: int, k: int):
        ref, srcs = self.pairs[k]
        ids = [ref] + list(srcs)
        p = pass_params(self.make_params, i, pass_)
        w, h = scaled_size(self.W, self.H, scale_size(self.rounds, i))
        imgs = np.stack([self.scaled_image(i, v) for v in ids])
        cams = self.cameras[ids].copy()
        if (w, h) != (self.W, self.H):
            sx, sy = f32(w) / f32(self.W), f32(h) / f32(self.H)
            for c in cams:
                K = c["K"]
                K[0] = f32(K[0]) * sx; K[2] = f32(K[2]) * sx; K[4] = f32(K[4]) * sy; K[5] = f32(K[5]) * sy
        cams["width"], cams["height"] = w, h
        p.depth_min = float(f32(cams[0]["depth_min"]) * f32(0.6))
        p.depth_max = float(f32(cams[0]["depth_max"]) * f32(1.2))
        depths = planes = views = states = None
        if p.geom_consistency:
            depths = np.stack([rescale_nearest(self.results[v]["depth"], w, h) for v in ids])
        if p.use_APD:
            states = np.ascontiguousarray(rescale_nearest(self.results[ref]["weak"], w, h))
        if p.state != FIRST_INIT:
            r = self.results[ref]
            d = rescale_nearest(r["depth"], w, h)
            n = rescale_nearest(r["normal"], w, h)
            planes = np.concatenate([n, d[..., None]], axis=-1).astype(f32)
            views = np.ascontiguousarray(rescale_nearest(r["views"], w, h))
        seed = self.seed + (i * 4 + pass_) * 65536 + k
        pl, st, vw = self.run_patchmatch(imgs, cams, p, depths, planes, views, states, seed)
        depth = pl[..., 3].copy()
        st = st.copy()
        bad = (depth < f32(p.depth_min)) | (depth > f32(p.depth_max))
        depth[bad] = 0
        st[bad] = UNKNOWN
        self.results[ref] = {"depth": depth, "normal": pl[..., :3].copy(), "weak": st, "views": vw.copy()}

    def run_pass(self, i: int, pass_: int, world: int = 1):
        """world == 1: the reference's order (every problem sees all earlier results of the same pass).
        world > 1: what `world` ranks with round-robin problem ownership compute - a rank sees its own earlier
        results of this pass and the other ranks' results of the previous pass (only their depth maps matter)."""
        if world == 1:
            for k in range(len(self.pairs)):
                self.process_problem(i, pass_, k)
            return
        start = list(self.results)
        merged = list(self.results)
        for r in range(world):
            self.results = list(start)
            for k in range(r, len(self.pairs), world):
                self.process_problem(i, pass_, k)
                merged[self.pairs[k][0]] = self.results[self.pairs[k][0]]
        self.results = merged

    def run(self):
        for i in range(self.rounds):
            for pass_ in range(4):
                self.run_pass(i, pass_)
