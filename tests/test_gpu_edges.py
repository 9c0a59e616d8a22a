"""Edge cases of the path, CUDA (through the C ABI) vs the oracle port, in both rounding contracts: the step
budget's ends, one-pixel frames, a camera inside the horizon, the general-domain fallback of the render loop (a
camera so far out that the branch-free div/sqrt domain does not apply; a horizon so large that a single RK4 stage
can jump inside r < Rs/2, where getGeodesicAcc returns zero, geodesics.h:33), and the measured kernel variants."""
import os

import numpy as np
import pytest

from parity import census

pytestmark = pytest.mark.gpu

CONTRACTS = [3, 7]   # strict, FMAD (flags: disk | dust [| RRT_FLAG_FMAD]); ORA_FLAG_* == RRT_FLAG_*


def pair(gpu, ora, sky_np, pos, yaw, pitch, w, h, flags, spin=0.99, **over):
    import relativisticraytracer_b200 as rrt
    import torch
    pg = rrt.default_params(spin_a=spin, flags=flags, **over)
    po = ora.default_params(spin_a=spin, flags=flags, **over)
    cg, co = rrt.camera_state_from(pos, yaw, pitch), ora.camera_from(pos, yaw, pitch)
    sky = gpu.create_sky(sky_np)
    planes = gpu.alloc_planes(w, h)
    gpu.read_counters(reset=True)
    out = gpu.render(pg, cg, rrt.effects_off(), sky, 1.0, w, h, planes=planes)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in planes.items()}
    g["rgba"], g["counters"] = out.cpu().numpy(), gpu.read_counters()
    sky.close()
    return ora.render(po, co, ora.effects_off(), sky_np, 1.0, w, h), g


def same_trajectories(f, g):
    for k in ("cls", "steps", "pos", "vel", "dir"):
        assert np.array_equal(g[k], getattr(f, k), equal_nan=True), k
    assert g["counters"] == f.counters


@pytest.mark.parametrize("flags", CONTRACTS)
@pytest.mark.parametrize("max_steps", [0, 1, 2, 37])
def test_step_budget_ends(gpu, ora, sky_smooth, flags, max_steps):
    f, g = pair(gpu, ora, sky_smooth, (15.0, 3.0, -30.0), -26.6, -5.1, 64, 36, flags, max_steps=max_steps)
    same_trajectories(f, g)
    assert int(g["steps"].max()) <= max_steps
    assert census(f, g)["rgb_frac_over_tol"] == 0.0


@pytest.mark.parametrize("flags", CONTRACTS)
@pytest.mark.parametrize("w,h", [(1, 1), (1, 7), (9, 1), (33, 5)])
def test_tiny_and_ragged_frames(gpu, ora, sky_smooth, flags, w, h):
    f, g = pair(gpu, ora, sky_smooth, (0.0, 10.0, -60.0), 0.0, -10.0, w, h, flags)
    same_trajectories(f, g)
    assert np.abs(g["rgba"].astype(int) - f.rgba.astype(int)).max() <= 1


@pytest.mark.parametrize("flags", CONTRACTS)
def test_camera_inside_the_horizon(gpu, ora, sky_smooth, flags):
    f, g = pair(gpu, ora, sky_smooth, (0.5, 1.0, 1.2), 30.0, 5.0, 40, 24, flags)
    same_trajectories(f, g)
    assert g["counters"]["n_captured"] == 40 * 24 and g["counters"]["rk4_steps"] == 0
    assert not g["rgba"][..., :3].any()          # black: T = 0 and nothing emitted (raymarcher.cu:47-51, 128)


@pytest.mark.parametrize("flags", CONTRACTS)
def test_far_camera_takes_the_general_path(gpu, ora, sky_smooth, flags):
    """|p|^2 >= 1e8: every step is redone with the guarded division / square root (rk4_step_general)"""
    f, g = pair(gpu, ora, sky_smooth, (0.0, 500.0, -20000.0), 0.0, -1.4, 48, 27, flags, max_steps=300)
    same_trajectories(f, g)
    assert g["counters"]["n_exhausted"] + g["counters"]["n_escaped"] == 48 * 27


@pytest.mark.parametrize("flags", CONTRACTS)
def test_stage_inside_half_horizon(gpu, ora, sky_smooth, flags):
    """event_horizon = 20 and a 30-unit step: RK4 stages land inside r < 10 = Rs/2, where the acceleration is
    defined as zero (geodesics.h:33) -- the branch-free loop must notice and redo those steps."""
    f, g = pair(gpu, ora, sky_smooth, (0.0, 2.0, -45.0), 0.0, -2.0, 64, 36, flags, event_horizon=20.0, step_size=100.0,
                disk_out=60.0, isco_radius=22.0, cloud_out=60.0)
    same_trajectories(f, g)
    assert g["counters"]["n_captured"] > 0 and g["counters"]["n_escaped"] + g["counters"]["n_exhausted"] > 0
    assert census(f, g)["rgb_frac_over_tol"] == 0.0


_VARIANT_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import relativisticraytracer_b200 as rrt
from parity import CAMERAS
w, h = 200, 117
r = rrt.Renderer(0)
sky = r.create_sky(rrt.procedural_sky(512, 256, seed=1234, stars=400))
planes = r.alloc_planes(w, h)
r.read_counters(reset=True)
out = r.render(rrt.default_params(spin_a=0.99, flags=3), rrt.camera_state_from(*CAMERAS["C1"]), rrt.default_effects(), sky, 1.0, w, h, planes=planes)
torch.cuda.synchronize()
res = {k: v.cpu().numpy() for k, v in planes.items()}
res["rgba"] = out.cpu().numpy()
cnt = r.read_counters()
res["counters"] = np.array([cnt[k] for k in sorted(cnt)], np.int64)
np.savez(sys.argv[2], **res)
"""


@pytest.mark.parametrize("variant", ["2", "3"])
def test_measured_kernel_variants_stay_bit_identical(gpu, sky_small, variant, tmp_path):
    """The two measured-and-rejected designs (packed f32x2; wavefront in a warp) are kept as evidence in a separate
    library (make -C csrc variants -> build/variants/librrt_b200_variants.so, RRT_KERNEL_VARIANT=2 / 3), not in the
    product; they must keep producing exactly the product kernel's strict-contract output (frame, planes, counters)."""
    import subprocess
    import sys
    import relativisticraytracer_b200 as rrt
    import torch
    from parity import CAMERAS
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "build", "variants", "librrt_b200_variants.so")
    if not os.path.exists(lib):
        pytest.skip("build/variants/librrt_b200_variants.so not built (make -C relativisticraytracer_b200/csrc variants)")
    w, h = 200, 117
    sky = gpu.create_sky(sky_small)
    planes = gpu.alloc_planes(w, h)
    gpu.read_counters(reset=True)
    out = gpu.render(rrt.default_params(spin_a=0.99, flags=3), rrt.camera_state_from(*CAMERAS["C1"]), rrt.default_effects(),
                     sky, 1.0, w, h, planes=planes)
    torch.cuda.synchronize()
    want = {k: v.cpu().numpy() for k, v in planes.items()}
    want["rgba"] = out.cpu().numpy()
    cnt = gpu.read_counters()
    want["counters"] = np.array([cnt[k] for k in sorted(cnt)], np.int64)
    sky.close()
    script, res = tmp_path / "variant.py", tmp_path / "variant.npz"
    script.write_text(_VARIANT_SCRIPT)
    env = dict(os.environ, RRT_B200_LIB=lib, RRT_KERNEL_VARIANT=variant)
    subprocess.run([sys.executable, str(script), root, str(res)], check=True, env=env, timeout=600)
    got = np.load(res)
    for k in want:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
