#!/usr/bin/env python
"""Golden vectors of the widened rows, made from the REFERENCE on a B200 (run under gpurun from the repo root):
  * whole pass schedule: oracle/ref_pipeline.py (the reference's driver loop) around oracle/_ref/libapd_ref.so (the reference's
    own APD.cu) on a 3-view 1010x64 scene (2 rounds x 4 passes) -> SHA-256 of every view's final depth / normal / state /
    selected-view maps plus a few sampled values;
  * fusion: the reference's unmodified RunFusion (oracle/_ref/libapd_fusion_ref.so) on those maps -> SHA-256 of the point
    list (coordinates and colours) and its length.
Writes gpurun_out/golden/pipeline_1010x64_v3.json (hashes) and pipeline_1010x64_v3_inputs.npz (8-bit images, cameras); copy both to
tests/golden/."""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fusion_tools as FT
import parity_tools as T
from apd_mvs_b200 import engine as E, pipeline as P
from apd_mvs_b200.scene import make_scene
from oracle import ref_pipeline as RP, ref_binding

W, H, V, S, SEED = 1010, 64, 3, 2, 9001


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_ref(imgs, cams, params, depths, planes, views, states, seed):
    ref = ref_binding.RefAPD(imgs, cams, T.clone_params(params), depths=depths, planes=planes, views=views, states=states, seed=seed)
    ref.run(); out = ref.outputs(); ref.close()
    return out


sc = make_scene(W, H, V - 1, device="cpu")
# 8-bit images, as the reference gets them from cv::imread; they are stored in the fixture (the generator's libm calls are
# not bit-reproducible across hosts)
img8 = np.clip(np.rint(sc["images"].numpy()), 0, 255).astype(np.uint8)
images, cams = img8.astype(np.float32), sc["cameras"]
pairs = P.ring_pairs(V, S)
rp = RP.RefPipeline(images, cams, pairs, E.default_params, run_ref, seed=SEED)
rp.run()
out = {"W": W, "H": H, "views": V, "src": S, "seed": SEED, "images_sha": sha(images), "results": []}
for v in range(V):
    r = rp.results[v]
    out["results"].append({"depth": sha(r["depth"]), "normal": sha(r["normal"]), "states": sha(r["weak"]), "views": sha(r["views"]),
                           "depth_samples": [float(x) for x in r["depth"][::16, ::101].ravel()[:12]]})
bgr = FT.colour_images(images)
with tempfile.TemporaryDirectory() as d:
    FT.write_dense_folder(d, list(range(V)), bgr, cams, [r["depth"] for r in rp.results], [r["normal"] for r in rp.results], [r["weak"] for r in rp.results])
    xyz, col = FT.run_reference_fusion(d, pairs)
out["fusion"] = {"points": int(len(xyz)), "xyz": sha(xyz), "bgr": sha(col)}
os.makedirs(os.path.join(ROOT, "gpurun_out", "golden"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "golden", "pipeline_1010x64_v3.json"), "w"), indent=1)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "golden", "pipeline_1010x64_v3_inputs.npz"), images=img8,
                    cameras=np.frombuffer(np.ascontiguousarray(cams).tobytes(), dtype=np.uint8))
print(json.dumps(out)[:400])
