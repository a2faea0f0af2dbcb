#!/usr/bin/env python
"""Generates the COLMAP fixture of tests/test_colmap_cpu.py and the golden outputs of the REFERENCE converter on it.

    python tests/golden/make_golden_colmap.py          (needs /root/reference/colmap2mvsnet.py; run in the authoring container)

Inputs  tests/golden/colmap_scene/{dslr_calibration_undistorted/{cameras,images,points3D}.{txt,bin}, images/*.png}
Golden  tests/golden/colmap_expected_{txt,bin}/{cams/%08d_cam.txt, pair.txt, images.json (sha256 of the written JPEGs)}
The reference script calls np.asscalar, which NumPy >= 1.23 no longer has; it is supplied here (test infrastructure only).
"""
import hashlib, importlib.util, json, os, shutil, struct, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCENE = os.path.join(HERE, "colmap_scene")


def build_scene():
    rng = np.random.default_rng(20231017)
    model_dir = os.path.join(SCENE, "dslr_calibration_undistorted")
    shutil.rmtree(SCENE, ignore_errors=True)
    os.makedirs(model_dir); os.makedirs(os.path.join(SCENE, "images"))
    # two cameras: PINHOLE (fx, fy, cx, cy) and SIMPLE_RADIAL (f, cx, cy, k)
    cams = {3: ("PINHOLE", 64, 48, [70.0, 71.5, 31.5, 23.25]), 7: ("SIMPLE_RADIAL", 60, 44, [66.25, 30.0, 22.0, 0.01])}
    n_img, n_pts = 9, 420
    pts = np.column_stack([rng.uniform(-2, 2, n_pts), rng.uniform(-1.5, 1.5, n_pts), rng.uniform(4, 7, n_pts)])
    pids = np.sort(rng.choice(np.arange(1, 5000), n_pts, replace=False))
    images = []
    ids = [2, 3, 5, 8, 9, 12, 13, 20, 21]            # COLMAP image ids: ascending, with gaps
    for k, iid in enumerate(ids):
        ang = 0.12 * (k - 4)
        if k == 6:
            ang = 0.12 * (5 - 4) + 0.0008          # nearly the same pose as image k = 5: triangulation angles below 1 degree
        ax = np.array([0.1 * np.sin(k), 1.0, 0.05 * np.cos(k)]); ax /= np.linalg.norm(ax)
        q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
        c = np.array([2.5 * np.sin(ang * 3), 0.1 * np.cos(k), 0.2 * k if k != 6 else 0.2 * 5 + 0.002])
        w, x, y, z = q
        R = np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                      [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                      [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])
        t = -R @ c
        cid = 3 if k % 3 else 7
        name, W, H, p = cams[cid]
        fx, fy, cx, cy = (p[0], p[0], p[1], p[2]) if name == "SIMPLE_RADIAL" else p
        Xc = pts @ R.T + t
        u, v = fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy
        vis = (Xc[:, 2] > 0) & (u > -20) & (u < W + 20) & (v > -20) & (v < H + 20) & (rng.random(n_pts) < 0.8)
        obs = [(float(u[j]), float(v[j]), int(pids[j])) for j in np.nonzero(vis)[0]]
        obs += [(float(rng.uniform(0, W)), float(rng.uniform(0, H)), -1) for _ in range(7)]      # untriangulated features
        order = rng.permutation(len(obs))
        images.append((iid, q, t, cid, "view_%02d.png" % iid, [obs[j] for j in order]))
    # ---- text model
    with open(os.path.join(model_dir, "cameras.txt"), "w") as f:
        f.write("# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n")
        for cid, (name, W, H, p) in cams.items():
            f.write(f"{cid} {name} {W} {H} " + " ".join(repr(float(x)) for x in p) + "\n")
    with open(os.path.join(model_dir, "images.txt"), "w") as f:
        f.write("# Image list with two lines of data per image:\n")
        for iid, q, t, cid, name, obs in images:
            f.write(f"{iid} " + " ".join(repr(float(x)) for x in list(q) + list(t)) + f" {cid} {name}\n")
            f.write(" ".join(f"{repr(u)} {repr(v)} {pid}" for u, v, pid in obs) + "\n")
    with open(os.path.join(model_dir, "points3D.txt"), "w") as f:
        f.write("# 3D point list with one line of data per point:\n")
        for pid, X in zip(pids, pts):
            f.write(f"{pid} {repr(float(X[0]))} {repr(float(X[1]))} {repr(float(X[2]))} 128 128 128 0.5 {images[0][0]} 0\n")
    # ---- binary model (COLMAP src/base/reconstruction.cc layouts)
    model_ids = {"PINHOLE": 1, "SIMPLE_RADIAL": 2}
    with open(os.path.join(model_dir, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(cams)))
        for cid, (name, W, H, p) in cams.items():
            f.write(struct.pack("<iiQQ", cid, model_ids[name], W, H) + struct.pack("<%dd" % len(p), *p))
    with open(os.path.join(model_dir, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(images)))
        for iid, q, t, cid, name, obs in images:
            f.write(struct.pack("<i7di", iid, *q, *t, cid) + name.encode() + b"\0" + struct.pack("<Q", len(obs)))
            for u, v, pid in obs:
                f.write(struct.pack("<ddq", u, v, pid))
    with open(os.path.join(model_dir, "points3D.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(pids)))
        for pid, X in zip(pids, pts):
            f.write(struct.pack("<Q3d3Bd", int(pid), *X, 128, 128, 128, 0.5) + struct.pack("<Q", 1) + struct.pack("<ii", images[0][0], 0))
    # ---- images of the two camera sizes
    import cv2
    for iid, q, t, cid, name, obs in images:
        _, W, H, _ = cams[cid]
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.stack([(xx * 3 + iid * 7) % 256, (yy * 5 + iid * 11) % 256, ((xx + yy) * 2 + iid) % 256], -1).astype(np.uint8)
        cv2.imwrite(os.path.join(SCENE, "images", name), img)


def run_reference(ext, out):
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: a.item()
    spec = importlib.util.spec_from_file_location("ref_colmap2mvsnet", "/root/reference/colmap2mvsnet.py")
    ref = importlib.util.module_from_spec(spec); sys.modules["ref_colmap2mvsnet"] = ref; spec.loader.exec_module(ref)     # importable by name: its process pool pickles calc_score
    shutil.rmtree(out, ignore_errors=True); os.makedirs(out)
    args = types.SimpleNamespace(dense_folder=SCENE, save_folder=out, max_d=192, interval_scale=1, scale_factor=1, theta0=5, sigma1=1, sigma2=10, model_ext=ext)
    ref.processing_single_scene(args)
    sha = {n: hashlib.sha256(open(os.path.join(out, "images", n), "rb").read()).hexdigest() for n in sorted(os.listdir(os.path.join(out, "images")))}
    shutil.rmtree(os.path.join(out, "images"))
    json.dump(sha, open(os.path.join(out, "images.json"), "w"), indent=1)


if __name__ == "__main__":
    build_scene()
    run_reference(".txt", os.path.join(HERE, "colmap_expected_txt"))
    run_reference(".bin", os.path.join(HERE, "colmap_expected_bin"))
    # a second golden with another scale factor and the inverse-depth rule (max_d = 0)
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: a.item()
    spec = importlib.util.spec_from_file_location("ref_colmap2mvsnet", "/root/reference/colmap2mvsnet.py")
    ref = importlib.util.module_from_spec(spec); sys.modules["ref_colmap2mvsnet"] = ref; spec.loader.exec_module(ref)
    out = os.path.join(HERE, "colmap_expected_scaled")
    shutil.rmtree(out, ignore_errors=True); os.makedirs(out)
    ref.processing_single_scene(types.SimpleNamespace(dense_folder=SCENE, save_folder=out, max_d=0, interval_scale=2.0, scale_factor=2.0, theta0=5, sigma1=1, sigma2=10, model_ext=".txt"))
    sha = {n: hashlib.sha256(open(os.path.join(out, "images", n), "rb").read()).hexdigest() for n in sorted(os.listdir(os.path.join(out, "images")))}
    shutil.rmtree(os.path.join(out, "images")); json.dump(sha, open(os.path.join(out, "images.json"), "w"), indent=1)
    print("golden written")
