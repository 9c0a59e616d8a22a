"""The split pipeline (csrc/rrt_split.cuh: trace / media / fold kernels over a sample pool) against the fused render_kernel.

A media sample (reference src/raymarcher.cu:67-115) does not feed back into the trajectory, so tracing first, evaluating
the samples of many rays densely packed and folding `I += e (1 - s) T; T *= s` per ray in step order must give the SAME
bits as evaluating every sample on the spot: uchar4 frames, every parity plane (hdr, dir, emis, pos, vel, cls, steps) and
every counter, in both rounding contracts, for every camera / medium combination, ragged sizes, bands, the step budget's
edges -- and for ANY pool size: a pool too small for the frame cuts it into passes, tiles that cannot get a slot are
traced again by the next pass, and whatever the enqueued passes leave is rendered by the closing fused sweep."""
import os

import numpy as np
import pytest

from parity import CAMERAS

pytestmark = pytest.mark.gpu


def _renderer_with_tracer(tracer):
    """RRT_TRACE (scalar | packed | auto) is read when a context is created."""
    import relativisticraytracer_b200 as rrt
    old = os.environ.get("RRT_TRACE")
    os.environ["RRT_TRACE"] = tracer
    try:
        return rrt.Renderer(0)
    finally:
        if old is None:
            os.environ.pop("RRT_TRACE", None)
        else:
            os.environ["RRT_TRACE"] = old


@pytest.fixture(scope="module", params=["packed", "scalar"])
def pair(built, request):
    """(fused, split) contexts; the split one with the packed f32x2 tracer (two rays per thread, the default under the FMAD
    contract; the strict contract has only the scalar tracer) or the scalar one."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import relativisticraytracer_b200 as rrt
    fused, split = rrt.Renderer(0), _renderer_with_tracer(request.param)
    fused.set_pipeline("fused")
    split.set_pipeline("split")
    yield fused, split, request.param
    fused.close()
    split.close()


def _frame(r, sky_np, cam, spin, flags, w, h, fx="default", band=None, time=1.0, **over):
    import relativisticraytracer_b200 as rrt
    import torch
    sky = r.create_sky(sky_np)
    planes = r.alloc_planes(w, h) if band is None else None
    r.read_counters(reset=True)
    kw = dict(band=band, layout=rrt.OUT_PACKED) if band is not None else {}
    out = r.render(rrt.default_params(spin_a=spin, flags=flags, **over), rrt.camera_state_from(*CAMERAS[cam]),
                   rrt.default_effects() if fx == "default" else rrt.effects_off(), sky, time, w, h, planes=planes, **kw)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in (planes or {}).items()}
    res["rgba"] = out.cpu().numpy()
    res["counters"] = r.read_counters()
    sky.close()
    return res


def _same(a, b, tag):
    assert a["counters"] == b["counters"], f"{tag}: {a['counters']} vs {b['counters']}"
    for k in a:
        if k != "counters":
            assert np.array_equal(a[k], b[k], equal_nan=True), f"{tag}: {k} differs on {(a[k] != b[k]).sum()} elements"


@pytest.mark.parametrize("cam", ["C0", "C1", "C2", "C3"])
@pytest.mark.parametrize("spin,flags", [(0.99, 7), (0.0, 7), (0.99, 5), (0.99, 6), (0.99, 3), (0.0, 1)])
def test_split_equals_fused(pair, sky_small, cam, spin, flags):
    fused, split, _ = pair
    w, h = 203, 117
    a, b = _frame(fused, sky_small, cam, spin, flags, w, h), _frame(split, sky_small, cam, spin, flags, w, h)
    _same(a, b, f"{cam} a={spin} flags={flags}")
    st = split.split_stats()
    assert st["passes_worked"] >= 1 and st["tiles_split"] >= 1 and st["tiles_split"] + st["tiles_swept"] >= st["tiles"], st
    if cam != "C2":
        assert a["counters"]["dense_samples"] > 0   # the comparison is about media, make sure there were some


def test_split_equals_fused_bands_budget_time_and_tiny_frames(pair, sky_small):
    import relativisticraytracer_b200 as rrt
    fused, split, _ = pair
    for band in (rrt.Band(0, 3, 8), rrt.Band(2, 3, 8), rrt.Band(1, 2, 1)):
        _same(_frame(fused, sky_small, "C1", 0.99, 7, 160, 90, band=band), _frame(split, sky_small, "C1", 0.99, 7, 160, 90, band=band), f"band {band.rank}/{band.nranks}")
    for steps in (0, 1, 8, 9, 333, 1999):
        _same(_frame(fused, sky_small, "C3", 0.99, 7, 64, 36, max_steps=steps), _frame(split, sky_small, "C3", 0.99, 7, 64, 36, max_steps=steps), f"max_steps={steps}")
    for t in (0.0, 7.25):
        _same(_frame(fused, sky_small, "C1", 0.99, 7, 96, 54, time=t), _frame(split, sky_small, "C1", 0.99, 7, 96, 54, time=t), f"time={t}")
    _same(_frame(fused, sky_small, "C3", 0.99, 7, 1, 1), _frame(split, sky_small, "C3", 0.99, 7, 1, 1), "1x1")
    _same(_frame(fused, sky_small, "C1", 0.99, 7, 3, 5, fx="off"), _frame(split, sky_small, "C1", 0.99, 7, 3, 5, fx="off"), "3x5")
    # non-default medium geometry (other zone radii, ring limits and step size)
    over = dict(disk_out=18.0, isco_radius=7.0, cloud_out=30.0, cloud_h=0.9, step_size=0.2)
    _same(_frame(fused, sky_small, "C1", 0.99, 7, 120, 67, **over), _frame(split, sky_small, "C1", 0.99, 7, 120, 67, **over), "non-default parameters")


@pytest.mark.parametrize("pool_kib,max_passes", [(32 * 1024, 32), (4 * 1024, 32), (4 * 1024, 2), (256, 3), (64, 1)])
def test_any_pool_size_gives_the_same_frame(built, sky_small, pair, pool_kib, max_passes):
    """Pools from 'several passes' down to 'smaller than one disk-plane tile's samples' (then every such tile gives up in
    every pass and the closing fused sweep renders it): passes, redo lists and the sweep must all lead to the same frame."""
    import relativisticraytracer_b200 as rrt
    fused, _, tracer = pair
    r = _renderer_with_tracer(tracer)
    try:
        r.set_pipeline("split")
        r.set_sample_pool(pool_kib * 1024, max_passes)
        want = _frame(fused, sky_small, "C3", 0.99, 7, 256, 144)
        worked = swept = 0
        for rep in range(3):   # the pass count adapts from frame to frame: every guess must give the same frame
            got = _frame(r, sky_small, "C3", 0.99, 7, 256, 144)
            _same(want, got, f"pool {pool_kib} KiB, {max_passes} passes, frame {rep}")
            st = r.split_stats()
            assert st["tiles_split"] + st["tiles_swept"] >= st["tiles"] > 0, st
            worked, swept = max(worked, st["passes_worked"]), max(swept, st["tiles_swept"])
        if pool_kib <= 4 * 1024:
            assert worked > 1 or swept > 0, "the small pool was expected to cut the frame into passes"
    finally:
        r.close()


def test_frames_in_flight_on_several_streams(pair, sky_small):
    """Frames on different streams own different pools; five streams share four pools (stream-ordered reuse)."""
    import torch
    import relativisticraytracer_b200 as rrt
    fused, split, _ = pair
    w, h = 160, 90
    sky_f, sky_s = fused.create_sky(sky_small), split.create_sky(sky_small)
    fx = rrt.default_effects()
    cams = ["C0", "C1", "C3", "C1", "C0", "C3", "C1"]
    want = [fused.render(rrt.default_params(spin_a=0.99, flags=7), rrt.camera_state_from(*CAMERAS[c]), fx, sky_f, 0.5 * i, w, h).cpu().numpy()
            for i, c in enumerate(cams)]
    streams = [torch.cuda.Stream() for _ in range(5)]
    outs = []
    split.set_frames_in_flight(4)
    try:
        for i, c in enumerate(cams):
            with torch.cuda.stream(streams[i % 5]):
                outs.append(split.render(rrt.default_params(spin_a=0.99, flags=7), rrt.camera_state_from(*CAMERAS[c]), fx, sky_s, 0.5 * i, w, h,
                                         stream=streams[i % 5]))
        torch.cuda.synchronize()
    finally:
        split.set_frames_in_flight(1)
    for i, o in enumerate(outs):
        assert np.array_equal(o.cpu().numpy(), want[i]), f"frame {i}"
    sky_f.close()
    sky_s.close()


@pytest.mark.parametrize("w,h,cam", [(1920, 1080, "C0"), (1920, 1080, "C3")])
def test_split_equals_fused_1080p(pair, sky_small, w, h, cam):
    fused, split, _ = pair
    _same(_frame(fused, sky_small, cam, 0.99, 7, w, h), _frame(split, sky_small, cam, 0.99, 7, w, h), f"{w}x{h} {cam}")
