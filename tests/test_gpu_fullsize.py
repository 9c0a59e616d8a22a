"""BASELINE.json's full sizes (1080p, 4K), where the CPU oracle would take minutes: size-independent properties
of the path instead of a pixel-by-pixel oracle comparison.

* a frame is a pure function of its inputs: two renders are identical, and a frame traced as N cyclic row bands
  (the multi-GPU partition) is identical to the single launch, pixel for pixel and counter for counter;
* the kernel's counters are consistent with its own planes (sum of per-pixel steps, class census);
* the total RK4 step count of the 4K bench frame is a known answer per rounding contract (a checksum of every
  trajectory's length: one flipped termination anywhere changes it)."""
import numpy as np
import pytest

from parity import CAMERAS

pytestmark = pytest.mark.gpu

# 3840x2160, a = 0.99, disk + dust, camera C0, reference default effects (lens distortion on), time 1.0
STEPS_4K = {"fmad": 8460399242, "strict": 8460399601}


def _render(gpu, sky, flags, w, h, cam="C0", spin=0.99, band=None, out=None, planes=None, layout=None):
    import relativisticraytracer_b200 as rrt
    kw = {}
    if layout is not None:
        kw["layout"] = layout
    return gpu.render(rrt.default_params(spin_a=spin, flags=flags), rrt.camera_state_from(*CAMERAS[cam]), rrt.default_effects(),
                      sky, 1.0, w, h, band=band, out=out, planes=planes, **kw)


@pytest.mark.parametrize("contract,flags", [("fmad", 7), ("strict", 3)])
def test_4k_frame_bands_counters_and_known_step_total(gpu, contract, flags):
    import relativisticraytracer_b200 as rrt
    import torch
    w, h = 3840, 2160
    sky = gpu.create_sky(rrt.procedural_sky(4096, 2048))
    planes = gpu.alloc_planes(w, h, names=("cls", "steps"))
    gpu.read_counters(reset=True)
    full = _render(gpu, sky, flags, w, h, planes=planes)
    torch.cuda.synchronize()
    cnt = gpu.read_counters(reset=True)
    assert cnt["rk4_steps"] == STEPS_4K[contract]
    assert int(planes["steps"].sum(dtype=torch.int64)) == cnt["rk4_steps"]
    cls = planes["cls"]
    assert cnt["n_captured"] == int(((cls & rrt.CLS_MASK) == rrt.CLS_CAPTURED).sum())
    assert cnt["n_exhausted"] == int(((cls & rrt.CLSF_EXHAUSTED) != 0).sum())
    assert cnt["n_touched"] == int(((cls & rrt.CLSF_TOUCHED) != 0).sum())
    assert cnt["n_captured"] + cnt["n_escaped"] + cnt["n_exhausted"] == w * h
    assert int(planes["steps"].max()) <= 2000
    # determinism
    again = _render(gpu, sky, flags, w, h)
    assert torch.equal(again, full)
    # 8 cyclic row bands (BASELINE config 4) == the single launch, and the counters add up the same
    frame = torch.zeros_like(full)
    for r in range(8):
        _render(gpu, sky, flags, w, h, band=rrt.Band(r, 8, 8), out=frame, layout=rrt.OUT_FRAME)
    torch.cuda.synchronize()
    cnt8 = gpu.read_counters(reset=True)
    assert torch.equal(frame, full)
    assert cnt8["rk4_steps"] == 2 * STEPS_4K[contract]          # the repeat render + the eight bands
    sky.close()


@pytest.mark.parametrize("flags,name", [(1 | 4, "config 2: disk only"), (3 | 4, "config 3: disk + dust")])
def test_1080p_frames_banded_equal_full(gpu, flags, name):
    import relativisticraytracer_b200 as rrt
    import torch
    w, h = 1920, 1080
    sky = gpu.create_sky(rrt.procedural_sky(4096, 2048))
    gpu.read_counters(reset=True)
    full = _render(gpu, sky, flags, w, h, cam="C1")
    torch.cuda.synchronize()
    c1 = gpu.read_counters(reset=True)
    frame = torch.zeros_like(full)
    for r in range(3):   # ragged: 1080 rows in groups of 16 over 3 ranks
        _render(gpu, sky, flags, w, h, cam="C1", band=rrt.Band(r, 3, 16), out=frame, layout=rrt.OUT_FRAME)
    torch.cuda.synchronize()
    c3 = gpu.read_counters(reset=True)
    assert torch.equal(frame, full), name
    assert c3 == c1, name
    if not (flags & 2):
        assert c1["dust_evals"] == 0 and c1["disk_evals"] > 0
    sky.close()
