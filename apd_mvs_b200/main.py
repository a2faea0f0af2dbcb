"""`python -m apd_mvs_b200.main dense_folder [gpu_index]` - the reference's command line (main.cpp:140-217) on the
scene layer: reads pair.txt, images/*.jpg, cams/*_cam.txt, runs the 4*round_num passes with everything resident on
the GPU, writes APD/<id>/{depths.dmb, normals.dmb, weak.bin, selected_views.bin} once at the end (the reference rewrites
them after every pass and deletes them after fusion) and fuses the depth maps on the GPU into APD/APD.ply (RunFusion,
APD.cpp:826-977 -> include/apd_fusion.h). JPEG decoding uses the Python cv2 module of this image (the C++ OpenCV the
reference links is absent)."""
from __future__ import annotations

import os
import sys

import numpy as np

from . import fusion as F
from . import io as IO
from . import pipeline as P
from .scene import CAMERA_DTYPE


def load_dense_folder(dense_folder: str):
    import cv2
    problems = IO.GenerateSampleList(os.path.join(dense_folder, "pair.txt"))
    ids = sorted({r for r, _ in problems} | {s for _, ss in problems for s in ss})
    index = {image_id: k for k, image_id in enumerate(ids)}
    images, cams = [], np.zeros(len(ids), dtype=CAMERA_DTYPE)
    for image_id in ids:
        name = IO.ToFormatIndex(image_id)
        img = cv2.imread(os.path.join(dense_folder, "images", name + ".jpg"), cv2.IMREAD_GRAYSCALE)     # APD.cpp:411
        if img is None:
            raise FileNotFoundError(f"images/{name}.jpg")
        images.append(img.astype(np.float32))                                                              # convertTo(CV_32FC1)
        cams[index[image_id]] = IO.ReadCamera(os.path.join(dense_folder, "cams", name + "_cam.txt"))
    shapes = {im.shape for im in images}
    if len(shapes) != 1:
        raise ValueError("Images may error, check it!")                                                   # CheckImages, main.cpp:51-70
    pairs = [(index[r], [index[s] for s in ss]) for r, ss in problems]
    return ids, np.stack(images), cams, pairs


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 1:
        print("USAGE: python -m apd_mvs_b200.main dense_folder [gpu_index]", file=sys.stderr)
        return 1
    dense_folder, gpu = argv[0], int(argv[1]) if len(argv) > 1 else 0
    ids, images, cams, pairs = load_dense_folder(dense_folder)
    print(f"There are {len(pairs)} problems needed to be processed!")
    scene = P.Scene(images, cams, pairs, device=gpu)
    print(f"Round nums: {scene.ComputeRoundNum()}")
    scene.Run()
    t = scene.Timing()
    print(f"PatchMatch GPU time {t['patchmatch_ms']:.1f} ms, wall {t['wall_ms']:.1f} ms, {t['launches']} kernel launches")
    for ref, _ in pairs:
        out = os.path.join(dense_folder, "APD", IO.ToFormatIndex(ids[ref]))
        os.makedirs(out, exist_ok=True)
        IO.WriteBinMat(os.path.join(out, "depths.dmb"), scene.Depth(ref))
        IO.WriteBinMat(os.path.join(out, "normals.dmb"), scene.Normal(ref))
        IO.WriteBinMat(os.path.join(out, "weak.bin"), scene.States(ref))
        IO.WriteBinMat(os.path.join(out, "selected_views.bin"), scene.SelectedViews(ref))
    # RunFusion (main.cpp:219): colour images at the depth-map size, optional blocks/mask_<id>.jpg
    import cv2
    fu = F.Fusion(len(ids), scene.W, scene.H, device=gpu)
    block_dir = os.path.join(dense_folder, "blocks")
    for k, image_id in enumerate(ids):
        if scene.ResultSize(k) != (scene.W, scene.H):
            raise RuntimeError("fusion needs full-resolution depth maps for every view (is every image a reference in pair.txt?)")
        bgr = cv2.imread(os.path.join(dense_folder, "images", IO.ToFormatIndex(image_id) + ".jpg"), cv2.IMREAD_COLOR)
        block = cv2.imread(os.path.join(block_dir, f"mask_{image_id}.jpg"), cv2.IMREAD_GRAYSCALE) if os.path.isdir(block_dir) else None
        fu.SetView(k, bgr, cams[k], scene.Depth(k), scene.Normal(k), scene.States(k), block)
    for ref, srcs in pairs:
        fu.AddProblem(ref, srcs)
    xyz, _ = fu.RunFusion()
    fu.ExportPointCloud(os.path.join(dense_folder, "APD", "APD.ply"))
    print(f"Fused {len(xyz)} points in {fu.Timing()['gpu_ms']:.1f} ms")
    fu.close()
    scene.close()
    print("All done")
    return 0


if __name__ == "__main__":
    sys.exit(main())
