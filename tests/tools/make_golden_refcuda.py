"""Golden uchar4 frames from the reference's OWN CUDA kernel (oracle/_ref/libref_cuda.so = src/raymarcher.cu,
unmodified, nvcc -O3 sm_100a with nvcc's default flags), rendered on a B200.  Run on the GPU box:
    python tests/tools/make_golden_refcuda.py gpurun_out/refcuda_frames.npz
and copy the file to tests/golden/.  Inputs are the ones tests/test_gpu_fmad.py re-creates: sky_small
(procedural_sky(512, 256, seed=1234, stars=400)), reference default CameraEffects, time 1.0, 160x90."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relativisticraytracer_b200 as rrt  # noqa: E402
from oracle import RefCuda  # noqa: E402
from parity import CAMERAS  # noqa: E402

W, H = 160, 90
sky = rrt.procedural_sky(512, 256, seed=1234, stars=400)
fx = rrt.default_effects()
ref = RefCuda()
out = {}
for cam in ("C0", "C1", "C2", "C3"):
    for spin, tag in ((0.0, "a000"), (0.99, "a099")):
        frame, _, _ = ref.render(spin, rrt.camera_state_from(*CAMERAS[cam]), fx, sky, 1.0, W, H)
        out[f"{cam}_{tag}"] = frame
np.savez_compressed(sys.argv[1], **out)
print("wrote", sys.argv[1], {k: v.shape for k, v in out.items()})
