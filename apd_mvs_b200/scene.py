"""Deterministic synthetic multi-view scenes — inputs for tests and bench.

World frame = reference camera frame. The surface is the lower envelope of three slanted planes
(z = z0 + a*x + b*y), i.e. the boundary of a union of half-spaces, so every view is rendered
analytically per pixel (first ray/plane crossing) with exact depth and no occlusion. Texture is
a sum of sinusoids in world (x, y); axis-aligned world rectangles are made textureless so that
WEAK regions exist for the deformation path.
"""
from __future__ import annotations

import math
import numpy as np
import torch

MASTER_SEED = 20231017
CURAND_SEED = 1234567

CAMERA_DTYPE = np.dtype([("K", "<f4", 9), ("R", "<f4", 9), ("t", "<f4", 3), ("c", "<f4", 3),
                         ("height", "<i4"), ("width", "<i4"), ("depth_min", "<f4"), ("depth_max", "<f4")])
assert CAMERA_DTYPE.itemsize == 112  # struct Camera, main.h:47-56

PLANES = np.array([[5.0, 0.30, 0.05], [5.0, -0.25, 0.10], [5.6, 0.02, -0.35]], dtype=np.float64)  # z0, a, b
Z_CENTRE = 5.0
DEPTH_MIN, DEPTH_MAX = 3.0, 8.0


def _look_at(c: np.ndarray, target: np.ndarray) -> np.ndarray:
    z = target - c
    z = z / np.linalg.norm(z)
    x = np.cross(np.array([0.0, 1.0, 0.0]), z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z])  # rows = camera axes in world coordinates


def make_cameras(W: int, H: int, n_src: int) -> np.ndarray:
    cams = np.zeros(n_src + 1, dtype=CAMERA_DTYPE)
    f = 0.8 * W
    K = np.array([f, 0, W / 2.0, 0, f, H / 2.0, 0, 0, 1], dtype=np.float64)
    target = np.array([0.0, 0.0, Z_CENTRE])
    for i in range(n_src + 1):
        if i == 0:
            c = np.zeros(3)
            R = np.eye(3)
        else:
            rad = 0.08 * Z_CENTRE * math.ceil(i / 8)
            ang = math.radians(45.0 * i + 10.0 * (math.ceil(i / 8) - 1))
            c = np.array([rad * math.cos(ang), rad * math.sin(ang), 0.0])
            R = _look_at(c, target)
        t = -R @ c
        cams[i]["K"] = K.astype(np.float32)
        cams[i]["R"] = R.reshape(-1).astype(np.float32)
        cams[i]["t"] = t.astype(np.float32)
        cams[i]["c"] = c.astype(np.float32)
        cams[i]["height"], cams[i]["width"] = H, W
        cams[i]["depth_min"], cams[i]["depth_max"] = DEPTH_MIN, DEPTH_MAX
    return cams


def _texture_params(rng: np.random.Generator, W: int, n_waves: int):
    px = Z_CENTRE / (0.8 * W)  # world size of one reference pixel at the scene centre
    lam = np.exp(rng.uniform(math.log(3.0), math.log(200.0), n_waves)) * px
    theta = rng.uniform(0, 2 * math.pi, n_waves)
    phase = rng.uniform(0, 2 * math.pi, n_waves)
    amp = rng.uniform(0.5, 1.0, n_waves)
    kx, ky = 2 * math.pi / lam * np.cos(theta), 2 * math.pi / lam * np.sin(theta)
    return kx, ky, phase, amp


def _rects(rng: np.random.Generator, n: int, frac: float):
    # world-space rectangles around the centre of the reference frustum
    half_w = Z_CENTRE * 0.5 / 0.8
    out = []
    for _ in range(n):
        cx, cy = rng.uniform(-0.8 * half_w, 0.8 * half_w, 2)
        sx, sy = rng.uniform(0.12, 0.3, 2) * half_w * math.sqrt(frac / 0.3)
        out.append((cx - sx, cx + sx, cy - sy, cy + sy, rng.uniform(60.0, 200.0)))
    return out


@torch.no_grad()
def make_scene(W: int, H: int, n_src: int, seed: int = MASTER_SEED, device: str = "cpu",
               textureless: bool = True, n_waves: int = 24, chunk_rows: int = 512, only=None, ref_index: int = 0) -> dict:
    """Returns images [N,H,W] float32 (0..255), cameras (CAMERA_DTYPE[N]), depth [N,H,W] float32
    (exact per-view depth), normal [H,W,3] float32 (view `ref_index`, world frame), weak_mask [H,W]
    bool (pixels of view `ref_index` on textureless rectangles). `only`: render just these views (the
    others stay zero) - a rank of the multi-GPU bench renders its own reference view's normal and mask
    and receives the images and depth maps from rank 0. The +-0.5 noise on the rectangles depends on
    the render order, so images are comparable only between calls with the same `only`."""
    rng = np.random.default_rng(seed)
    cams = make_cameras(W, H, n_src)
    kx, ky, phase, amp = _texture_params(rng, W, n_waves)
    rects = _rects(rng, 6, 0.3) if textureless else []
    dev = torch.device(device)
    dd = torch.float64
    kx_t, ky_t = torch.tensor(kx, dtype=dd, device=dev), torch.tensor(ky, dtype=dd, device=dev)
    ph_t, am_t = torch.tensor(phase, dtype=dd, device=dev), torch.tensor(amp, dtype=dd, device=dev)
    planes = torch.tensor(PLANES, dtype=dd, device=dev)
    N = n_src + 1
    images = torch.empty((N, H, W), dtype=torch.float32, device=dev)
    depth = torch.empty((N, H, W), dtype=torch.float32, device=dev)
    normal = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    weak = torch.zeros((H, W), dtype=torch.bool, device=dev)
    noise_gen = torch.Generator(device="cpu").manual_seed(seed + 17)
    xs = torch.arange(W, dtype=dd, device=dev)
    if only is not None:
        images.zero_(); depth.zero_()
    for i in (range(N) if only is None else only):
        K = cams[i]["K"].astype(np.float64)
        R = torch.tensor(cams[i]["R"].astype(np.float64).reshape(3, 3), device=dev)
        c = torch.tensor(cams[i]["c"].astype(np.float64), device=dev)
        for r0 in range(0, H, chunk_rows):
            r1 = min(H, r0 + chunk_rows)
            ys = torch.arange(r0, r1, dtype=dd, device=dev)
            dxc = ((xs - K[2]) / K[0])[None, :].expand(r1 - r0, W)
            dyc = ((ys - K[5]) / K[4])[:, None].expand(r1 - r0, W)
            dcam = torch.stack([dxc, dyc, torch.ones_like(dxc)], dim=-1)      # camera-frame ray, z = 1
            dw = dcam @ R                                                        # world ray = R^T d
            # plane i: z - a x - b y - z0 = 0  ->  t = (z0 + a cx + b cy - cz) / (dz - a dx - b dy)
            best_t = torch.full((r1 - r0, W), float("inf"), dtype=dd, device=dev)
            best_i = torch.zeros((r1 - r0, W), dtype=torch.long, device=dev)
            for pi in range(planes.shape[0]):
                z0, a, b = planes[pi]
                num = z0 + a * c[0] + b * c[1] - c[2]
                den = dw[..., 2] - a * dw[..., 0] - b * dw[..., 1]
                t = num / den
                ok = (t > 0) & (t < best_t)
                best_t = torch.where(ok, t, best_t)
                best_i = torch.where(ok, torch.full_like(best_i, pi), best_i)
            X = c[None, None, :] + best_t[..., None] * dw
            val = torch.zeros((r1 - r0, W), dtype=dd, device=dev)
            for k in range(n_waves):
                val += am_t[k] * torch.sin(kx_t[k] * X[..., 0] + ky_t[k] * X[..., 1] + ph_t[k] + 0.7 * best_i)
            val = 127.5 + 110.0 * val / float(np.sqrt((amp ** 2).sum() / 2.0) * 2.2)
            val = val.clamp(0.0, 255.0)
            for (xa, xb, ya, yb, level) in rects:
                inside = (X[..., 0] > xa) & (X[..., 0] < xb) & (X[..., 1] > ya) & (X[..., 1] < yb)
                nz = (torch.rand((r1 - r0, W), generator=noise_gen, dtype=torch.float32) - 0.5).to(dev, dd)
                val = torch.where(inside, level + nz, val)
                if i == ref_index:
                    weak[r0:r1] |= inside
            images[i, r0:r1] = val.to(torch.float32)
            depth[i, r0:r1] = best_t.to(torch.float32)        # camera-frame z of the hit (ray has z = 1)
            if i == ref_index:
                n = torch.stack([planes[best_i, 1], planes[best_i, 2], -torch.ones_like(best_t)], dim=-1)
                n = n / n.norm(dim=-1, keepdim=True)
                normal[r0:r1] = n.to(torch.float32)
    return {"images": images, "cameras": cams, "depth": depth, "normal": normal, "weak_mask": weak,
            "W": W, "H": H, "n_src": n_src, "ref_index": ref_index}


def make_priors(scene: dict, seed: int = MASTER_SEED + 1, depth_noise: float = 0.01, normal_deg: float = 5.0,
                views_mask: int = 0b1111, order=None):
    """Priors of a refinement pass (SURVEY §8d cfg 3): noisy GT depth/normal, WEAK on the textureless
    rectangles, UNKNOWN on the 6-px border, STRONG elsewhere; noisy GT depth for the source views.
    `order`: view indices [reference, src_1 .. src_S] of this problem inside scene["depth"] (default: all, in order);
    order[0] must be the view scene["normal"] / scene["weak_mask"] were rendered for."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    H, W = scene["H"], scene["W"]
    depth = scene["depth"].cpu()
    if order is not None:
        assert order[0] == scene.get("ref_index", 0)
        depth = depth[list(order)]
    nrm = scene["normal"].cpu()
    d0 = depth[0] * (1.0 + depth_noise * torch.randn((H, W), generator=g))
    pert = torch.randn((H, W, 3), generator=g) * math.tan(math.radians(normal_deg)) / math.sqrt(2.0)
    n = nrm + pert
    n = n / n.norm(dim=-1, keepdim=True)
    planes = torch.cat([n, d0[..., None]], dim=-1).to(torch.float32).contiguous().numpy()
    states = np.full((H, W), 1, dtype=np.uint8)
    states[scene["weak_mask"].cpu().numpy()] = 0
    states[:6, :] = 2; states[-6:, :] = 2; states[:, :6] = 2; states[:, -6:] = 2
    n_src = depth.shape[0] - 1
    views = np.full((H, W), views_mask & ((1 << n_src) - 1), dtype=np.uint32)
    src_depth = (depth * (1.0 + depth_noise * torch.randn(depth.shape, generator=g))).to(torch.float32).contiguous().numpy()
    return {"planes": planes, "states": states, "views": views, "depths": src_depth}
