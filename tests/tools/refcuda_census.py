"""Byte-level census of our uchar4 frames against the reference's own CUDA kernel (oracle/_ref/libref_cuda.so,
unmodified src/raymarcher.cu built with nvcc defaults) on the same GPU, for both rounding contracts.
Usage: python tests/tools/refcuda_census.py [w h]   (GPU box only; test infrastructure)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relativisticraytracer_b200 as rrt  # noqa: E402
from oracle import RefCuda  # noqa: E402
from parity import CAMERAS  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
r = rrt.Renderer(0)
sky_np = rrt.procedural_sky(4096, 2048)
sky = r.create_sky(sky_np)
ref = RefCuda()
fx = rrt.default_effects()
rows = []
for cam in ("C0", "C1", "C2", "C3"):
    for spin in (0.0, 0.99):
        cg = rrt.camera_state_from(*CAMERAS[cam])
        ref_rgba, _, _ = ref.render(spin, cg, fx, sky_np, 1.0, w, h)
        row = {"camera": cam, "spin": spin, "pixels": w * h}
        for name, flags in (("fmad", 7), ("strict", 3)):
            planes = r.alloc_planes(w, h, names=("cls",))
            out = r.render(rrt.default_params(spin_a=spin, flags=flags), cg, fx, sky, 1.0, w, h, planes=planes)
            torch.cuda.synchronize()
            ours = out.cpu().numpy()
            cls = planes["cls"].cpu().numpy()[::-1]
            touched = (cls & rrt.CLSF_TOUCHED) != 0
            d = np.abs(ours.astype(int) - ref_rgba.astype(int)).max(axis=-1)
            row[name] = {"touched_pixels": int(touched.sum()),
                         "untouched_differing": int((d[~touched] > 0).sum()), "untouched_max": int(d[~touched].max()),
                         "touched_differing": int((d[touched] > 0).sum()), "touched_max": int(d[touched].max()) if touched.any() else 0}
        rows.append(row)
        print(json.dumps(row), flush=True)
