"""COLMAP sparse model -> the dense_folder layout the pipeline reads (cams/%08d_cam.txt, pair.txt, images/%08d.jpg).

Replaces the reference's `colmap2mvsnet.py` (:280-469 `calc_score` / `processing_single_scene`; :58-277 are COLMAP's model
readers). Same command line, same output files. Not a translation:

* the model is read into ARRAYS (quaternions / translations [N, .], a CSR list of observed 3-D point ids per image, a
  point table), not into dicts of named tuples of small arrays;
* the view-selection score - for every image pair the number of shared 3-D points, zeroed when the 75th percentile of the
  triangulation angles is below 1 degree (reference :280-303, an O(N^2 x points) Python double loop over a process pool)
  - is a handful of tensor operations: the shared counts are ONE matrix product of the image x point incidence matrices,
  the angle percentile a masked sort per reference image; they run on the GPU when there is one (`--device`);
* depth ranges are vectorised per image.

    python -m apd_mvs_b200.colmap2mvsnet --dense_folder <scene> --save_folder <out> [--model_ext .txt|.bin] [--max_d 192]
        [--interval_scale 1] [--scale_factor 1] [--device cuda:0|cpu]

Golden outputs of the reference script on a synthetic model are committed under tests/golden/colmap_* (generator:
tests/golden/make_golden_colmap.py); tests/test_colmap_cpu.py demands byte-identical cams/ and pair.txt.
"""
from __future__ import annotations

import argparse
import os
import shutil
import struct
from dataclasses import dataclass, field

import numpy as np

# model id -> (name, number of parameters), COLMAP src/base/camera_models.h; the converter only needs fx, fy, cx, cy
CAMERA_MODELS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
                 5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
                 9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12)}
SINGLE_FOCAL = {"SIMPLE_PINHOLE", "SIMPLE_RADIAL", "SIMPLE_RADIAL_FISHEYE", "RADIAL", "RADIAL_FISHEYE"}     # params start f, cx, cy


@dataclass
class SparseModel:
    cameras: dict = field(default_factory=dict)            # camera id -> (model name, width, height, params float64[])
    image_ids: np.ndarray = None                           # [N] ascending COLMAP image ids (the reference renumbers them 1..N in this order)
    qvec: np.ndarray = None                                # [N, 4] (w, x, y, z)
    tvec: np.ndarray = None                                # [N, 3]
    camera_id: np.ndarray = None                           # [N]
    names: list = None                                     # [N]
    obs_ptr: np.ndarray = None                             # [N + 1] CSR offsets into obs_pid
    obs_pid: np.ndarray = None                             # 3-D point id of every 2-D observation, -1 = none
    point_ids: np.ndarray = None                           # [P] ascending
    point_xyz: np.ndarray = None                           # [P, 3]


def _finish(cams, recs, pts) -> SparseModel:
    recs.sort(key=lambda r: r[0])
    m = SparseModel(cameras=cams)
    m.image_ids = np.array([r[0] for r in recs], np.int64)
    m.qvec = np.array([r[1] for r in recs], np.float64).reshape(-1, 4)
    m.tvec = np.array([r[2] for r in recs], np.float64).reshape(-1, 3)
    m.camera_id = np.array([r[3] for r in recs], np.int64)
    m.names = [r[4] for r in recs]
    m.obs_ptr = np.zeros(len(recs) + 1, np.int64)
    np.cumsum([len(r[5]) for r in recs], out=m.obs_ptr[1:])
    m.obs_pid = np.concatenate([r[5] for r in recs]) if recs else np.zeros(0, np.int64)
    order = np.argsort(pts[0], kind="stable")
    m.point_ids, m.point_xyz = pts[0][order], pts[1][order]
    return m


def read_model_text(folder: str) -> SparseModel:
    cams, recs = {}, []
    for ln in open(os.path.join(folder, "cameras.txt")):
        f = ln.split()
        if f and not f[0].startswith("#"):
            cams[int(f[0])] = (f[1], int(f[2]), int(f[3]), np.array(f[4:], np.float64))
    lines = [ln for ln in open(os.path.join(folder, "images.txt")) if not ln.lstrip().startswith("#")]
    k = 0
    while k < len(lines):                  # two lines per image; the second (observations) may be empty
        f = lines[k].split()
        if not f:
            k += 1
            continue
        obs = lines[k + 1].split() if k + 1 < len(lines) else []
        recs.append((int(f[0]), [float(x) for x in f[1:5]], [float(x) for x in f[5:8]], int(f[8]), f[9], np.array(obs[2::3], np.int64)))
        k += 2
    ids, xyz = [], []
    for ln in open(os.path.join(folder, "points3D.txt")):
        f = ln.split()
        if f and not f[0].startswith("#"):
            ids.append(int(f[0])); xyz.append([float(f[1]), float(f[2]), float(f[3])])
    return _finish(cams, recs, (np.array(ids, np.int64), np.array(xyz, np.float64).reshape(-1, 3)))


def read_model_binary(folder: str) -> SparseModel:
    cams, recs = {}, []
    buf = open(os.path.join(folder, "cameras.bin"), "rb").read()
    n, = struct.unpack_from("<Q", buf, 0); o = 8
    for _ in range(n):
        cid, mid, w, h = struct.unpack_from("<iiQQ", buf, o); o += 24
        name, npar = CAMERA_MODELS[mid]
        cams[cid] = (name, w, h, np.frombuffer(buf, "<f8", npar, o).copy()); o += 8 * npar
    buf = open(os.path.join(folder, "images.bin"), "rb").read()
    n, = struct.unpack_from("<Q", buf, 0); o = 8
    obs_t = np.dtype([("x", "<f8"), ("y", "<f8"), ("pid", "<i8")])
    for _ in range(n):
        iid = struct.unpack_from("<i", buf, o)[0]
        q = struct.unpack_from("<4d", buf, o + 4); t = struct.unpack_from("<3d", buf, o + 36)
        cid = struct.unpack_from("<i", buf, o + 60)[0]; o += 64
        e = buf.index(b"\0", o); name = buf[o:e].decode("utf-8"); o = e + 1
        m, = struct.unpack_from("<Q", buf, o); o += 8
        recs.append((iid, list(q), list(t), cid, name, np.frombuffer(buf, obs_t, m, o)["pid"].astype(np.int64))); o += 24 * m
    buf = open(os.path.join(folder, "points3D.bin"), "rb").read()
    n, = struct.unpack_from("<Q", buf, 0); o = 8
    ids, xyz = np.empty(n, np.int64), np.empty((n, 3), np.float64)
    for k in range(n):
        ids[k] = struct.unpack_from("<Q", buf, o)[0]; xyz[k] = struct.unpack_from("<3d", buf, o + 8); o += 43
        track, = struct.unpack_from("<Q", buf, o); o += 8 + 8 * track
    return _finish(cams, recs, (ids, xyz))


def read_model(folder: str, ext: str) -> SparseModel:
    return read_model_text(folder) if ext == ".txt" else read_model_binary(folder)


def rotations(qvec: np.ndarray) -> np.ndarray:
    """[N, 4] (w, x, y, z) -> [N, 3, 3]; element expressions as in the reference's qvec2rotmat (so that str() of every entry agrees)."""
    w, x, y, z = qvec[:, 0], qvec[:, 1], qvec[:, 2], qvec[:, 3]
    R = np.empty((len(qvec), 3, 3), np.float64)
    R[:, 0, 0] = 1 - 2 * y ** 2 - 2 * z ** 2; R[:, 0, 1] = 2 * x * y - 2 * w * z; R[:, 0, 2] = 2 * z * x + 2 * w * y
    R[:, 1, 0] = 2 * x * y + 2 * w * z; R[:, 1, 1] = 1 - 2 * x ** 2 - 2 * z ** 2; R[:, 1, 2] = 2 * y * z - 2 * w * x
    R[:, 2, 0] = 2 * z * x - 2 * w * y; R[:, 2, 1] = 2 * y * z + 2 * w * x; R[:, 2, 2] = 1 - 2 * x ** 2 - 2 * y ** 2
    return R


def intrinsics(model: SparseModel, scale_factor: float) -> dict:
    out = {}
    for cid, (name, _w, _h, p) in model.cameras.items():
        fx, fy, cx, cy = (p[0], p[0], p[1], p[2]) if name in SINGLE_FOCAL else (p[0], p[1], p[2], p[3])
        out[cid] = np.array([[fx / scale_factor, 0, cx / scale_factor], [0, fy / scale_factor, cy / scale_factor], [0, 0, 1]])
    return out


def _dense_points(model: SparseModel):
    """Observed point ids -> row indices of the point table (-1 stays -1; ids missing from points3D are an error, as in the reference)."""
    idx = np.searchsorted(model.point_ids, np.maximum(model.obs_pid, 0))
    idx = np.minimum(idx, max(len(model.point_ids) - 1, 0))
    ok = model.obs_pid >= 0
    if ok.any() and not np.array_equal(model.point_ids[idx[ok]], model.obs_pid[ok]):
        raise KeyError("an image observes a 3-D point id that points3D does not contain")
    return np.where(ok, idx, -1)


def depth_ranges(model: SparseModel, R: np.ndarray, K: dict, max_d: int, interval_scale: float):
    """Reference :363-396: per image the 1 % / 99 % depth quantiles of its triangulated points, relaxed by 0.75 / 1.25."""
    dense = _dense_points(model)
    out = []
    for i in range(len(model.image_ids)):
        p = dense[model.obs_ptr[i]:model.obs_ptr[i + 1]]
        X = model.point_xyz[p[p >= 0]]
        depth_min = depth_max = 0
        if len(X):
            z = np.sort(((R[i, 2, 0] * X[:, 0] + R[i, 2, 1] * X[:, 1]) + R[i, 2, 2] * X[:, 2]) + model.tvec[i, 2])
            depth_min, depth_max = z[int(len(z) * .01)] * 0.75, z[int(len(z) * .99)] * 1.25
        if max_d == 0:     # inverse-depth sampling: one pixel of disparity at depth_min (reference :382-392)
            Ki = K[int(model.camera_id[i])]
            Kinv, Rinv = np.linalg.inv(Ki), np.linalg.inv(R[i])
            P1 = Rinv @ (Kinv @ [Ki[0, 2], Ki[1, 2], 1] * depth_min - model.tvec[i])
            P2 = Rinv @ (Kinv @ [Ki[0, 2] + 1, Ki[1, 2], 1] * depth_min - model.tvec[i])
            depth_num = (1 / depth_min - 1 / depth_max) / (1 / depth_min - 1 / (depth_min + np.linalg.norm(P2 - P1)))
        else:
            depth_num = max_d
        out.append((depth_min, (depth_max - depth_min) / (depth_num - 1) / interval_scale, depth_num, depth_max))
    return out


def view_selection_scores(model: SparseModel, R: np.ndarray, device: str = "cpu") -> np.ndarray:
    """score[i, j] = number of 3-D points images i and j share (observations of i counted with their multiplicity, i < j,
    mirrored), 0 if the 75th percentile of the triangulation angles at those points is below 1 degree (reference :280-303)."""
    import torch
    dev = torch.device(device)
    N, P = len(model.image_ids), len(model.point_ids)
    dense = torch.from_numpy(_dense_points(model)).to(dev)
    ptr = model.obs_ptr
    X = torch.from_numpy(model.point_xyz).to(dev)
    Rt = torch.from_numpy(R).to(dev); t = torch.from_numpy(model.tvec).to(dev)
    C = -(Rt.transpose(1, 2) @ t[:, :, None])[:, :, 0]                                   # camera centres [N, 3]
    count = torch.zeros((N, P), dtype=torch.float32, device=dev)                         # observations of point p in image i
    for i in range(N):
        p = dense[ptr[i]:ptr[i + 1]]
        p = p[p >= 0]
        count[i].index_add_(0, p, torch.ones(len(p), dtype=torch.float32, device=dev))
    seen = (count > 0).to(torch.float32)
    shared = count @ seen.T                                                              # exact: small integers in fp32
    score = torch.zeros((N, N), dtype=torch.float64, device=dev)
    for i in range(N - 1):
        p = dense[ptr[i]:ptr[i + 1]]
        p = p[p >= 0]                                                                    # with multiplicity, as the reference iterates id_i
        if len(p) == 0:
            continue
        a = C[i][None, :] - X[p]                                                         # [m, 3]
        b = C[i + 1:, None, :] - X[p][None, :, :]                                        # [N-i-1, m, 3]
        dot = (a[None] * b).sum(-1)
        cos = dot / a.norm(dim=-1)[None] / b.norm(dim=-1)
        theta = (180 / np.pi) * torch.arccos(cos)
        mask = seen[i + 1:][:, p] > 0
        theta = torch.where(mask, theta, torch.full_like(theta, float("inf")))
        srt, _ = torch.sort(theta, dim=1)
        n = mask.sum(1)
        k = torch.clamp((n.to(torch.float64) * 0.75).to(torch.int64), max=srt.shape[1] - 1)
        tri = srt.gather(1, k[:, None])[:, 0]
        s = shared[i, i + 1:].to(torch.float64)
        s = torch.where((n > 0) & (tri < 1), torch.zeros_like(s), s)
        score[i, i + 1:] = s
        score[i + 1:, i] = s
    return score.cpu().numpy()


def select_views(score: np.ndarray):
    """Reference :412-416: the (at most) 20 best-scoring partners of every image, in np.argsort's order."""
    num_view = min(20, len(score) - 1)
    return [[(int(k), score[i, k]) for k in np.argsort(score[i])[::-1][:num_view]] for i in range(len(score))]


def write_cams(cam_dir, model, R, K, ranges):
    os.makedirs(cam_dir, exist_ok=True)
    for i in range(len(model.image_ids)):
        E = np.zeros((4, 4)); E[:3, :3] = R[i]; E[:3, 3] = model.tvec[i]; E[3, 3] = 1
        Ki = K[int(model.camera_id[i])]
        with open(os.path.join(cam_dir, "%08d_cam.txt" % i), "w") as f:
            f.write("extrinsic\n" + "".join("".join(str(E[j, k]) + " " for k in range(4)) + "\n" for j in range(4)))
            f.write("\nintrinsic\n" + "".join("".join(str(Ki[j, k]) + " " for k in range(3)) + "\n" for j in range(3)))
            f.write("\n%f %f %f %f\n" % ranges[i])


def write_pairs(path, view_sel):
    with open(path, "w") as f:
        f.write("%d\n" % len(view_sel))
        for i, sel in enumerate(view_sel):
            f.write("%d\n%d " % (i, len(sel)) + "".join("%d %d " % (k, s) for k, s in sel) + "\n")


def convert_images(model, image_dir, out_dir, scale_factor):
    """Reference :441-459: pad every image to the largest size, nearest-neighbour down-scale, re-encode as %08d.jpg."""
    import cv2
    imgs = [cv2.imread(os.path.join(image_dir, n)) for n in model.names]
    H, W = max(im.shape[0] for im in imgs), max(im.shape[1] for im in imgs)
    for i, im in enumerate(imgs):
        im = np.pad(im, ((0, H - im.shape[0]), (0, W - im.shape[1]), (0, 0)), "constant")
        im = cv2.resize(im, (int(im.shape[1] / scale_factor), int(im.shape[0] / scale_factor)), interpolation=cv2.INTER_NEAREST)
        cv2.imwrite(os.path.join(out_dir, "%08d.jpg" % i), im)


def processing_single_scene(args):
    image_dir = os.path.join(args.dense_folder, "images")
    model_dir = os.path.join(args.dense_folder, "dslr_calibration_undistorted")
    cam_dir, image_out = os.path.join(args.save_folder, "cams"), os.path.join(args.save_folder, "images")
    for d in (image_out, cam_dir):
        if os.path.exists(d):
            shutil.rmtree(d)
    os.makedirs(image_out)
    model = read_model(model_dir, args.model_ext)
    R = rotations(model.qvec)
    K = intrinsics(model, args.scale_factor)
    ranges = depth_ranges(model, R, K, args.max_d, args.interval_scale)
    device = args.device
    if device is None:
        import torch
        device = "cuda:0" if torch.cuda.is_available() else "cpu"
    score = view_selection_scores(model, R, device)
    write_cams(cam_dir, model, R, K, ranges)
    write_pairs(os.path.join(args.save_folder, "pair.txt"), select_views(score))
    convert_images(model, image_dir, image_out, args.scale_factor)
    return score


def main(argv=None):
    ap = argparse.ArgumentParser(description="Convert colmap camera")
    ap.add_argument("--dense_folder", required=True, type=str)
    ap.add_argument("--save_folder", required=True, type=str)
    ap.add_argument("--max_d", type=int, default=192)
    ap.add_argument("--interval_scale", type=float, default=1)
    ap.add_argument("--scale_factor", type=float, default=1)
    ap.add_argument("--theta0", type=float, default=5)          # accepted and unused, as in the reference (:296 is commented out)
    ap.add_argument("--sigma1", type=float, default=1)
    ap.add_argument("--sigma2", type=float, default=10)
    ap.add_argument("--model_ext", type=str, default=".txt", choices=[".txt", ".bin"])
    ap.add_argument("--device", type=str, default=None, help="torch device of the view-selection scoring (default: cuda:0 if present)")
    args = ap.parse_args(argv)
    os.makedirs(args.save_folder, exist_ok=True)
    processing_single_scene(args)


if __name__ == "__main__":
    main()
