"""Frame sequences through the C ABI: frames in flight on several streams (rrt_render_host_async,
FramePipeline) and the frame-parallel path renderer with its sink give exactly the frames of the one-at-a-time
calls.  Every pixel is a pure function of its frame's inputs, so equality is bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 96, 54


def _setup(gpu, sky_np):
    import relativisticraytracer_b200 as rrt
    sky = gpu.create_sky(sky_np)
    prm = rrt.default_params(spin_a=0.99)
    fx = rrt.default_effects()
    return rrt, sky, prm, fx


def test_render_host_async_slots(gpu, sky_small):
    import torch
    rrt, sky, prm, fx = _setup(gpu, sky_small)
    cam = rrt.camera_state_from((15.0, 3.0, -30.0), -26.6, -5.1)
    want = []
    for t in (0.5, 1.0, 1.5, 2.0):
        out = np.zeros((H, W, 4), np.uint8)
        gpu.render_host(prm, cam, fx, sky, t, W, H, out)
        want.append(out)
    streams = [torch.cuda.Stream() for _ in range(4)]
    hosts = [torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(4)]
    for k, t in enumerate((0.5, 1.0, 1.5, 2.0)):
        gpu.render_host_async(prm, cam, fx, sky, t, W, H, hosts[k], slot=k, stream=streams[k])
    torch.cuda.synchronize()
    for k in range(4):
        assert np.array_equal(hosts[k].numpy(), want[k]), k
    with pytest.raises(rrt.RrtError):
        gpu.render_host_async(prm, cam, fx, sky, 1.0, W, H, hosts[0], slot=rrt.HOST_SLOTS, stream=streams[0])
    assert not np.array_equal(want[0], want[3])   # the disk moves with time: the frames really differ


@pytest.mark.parametrize("to_host", [False, True])
def test_frame_pipeline_equals_one_at_a_time(gpu, sky_small, to_host):
    import torch
    from relativisticraytracer_b200.parallel import FramePipeline
    rrt, sky, prm, fx = _setup(gpu, sky_small)
    cam = rrt.camera_state_from((0.0, 10.0, -60.0), 0.0, -10.0)
    times = [1.0 + 0.25 * k for k in range(5)]
    want = [gpu.render(prm, cam, fx, sky, t, W, H).cpu().numpy() for t in times]
    pipe = FramePipeline(gpu, W, H, depth=2, to_host=to_host)
    pipe.begin()
    got = []
    for t in times:
        pipe.submit(prm, cam, fx, sky, t)
        pipe.end()
        torch.cuda.synchronize()
        f = pipe.last_frame()
        got.append(f.numpy().copy() if to_host else f.cpu().numpy())
    for k in range(len(times)):
        assert np.array_equal(got[k], want[k]), k
    # sharing the resident-CTA slots between the frames in flight changes the schedule, not the pixels
    gpu.set_frames_in_flight(2)
    try:
        pipe3 = FramePipeline(gpu, W, H, depth=4, to_host=to_host)
        pipe3.begin()
        for t in times:
            pipe3.submit(prm, cam, fx, sky, t)
        pipe3.end()
        torch.cuda.synchronize()
        f = pipe3.last_frame()
        assert np.array_equal(f.numpy() if to_host else f.cpu().numpy(), want[-1])
    finally:
        gpu.set_frames_in_flight(1)
    with pytest.raises(rrt.RrtError):
        gpu.set_frames_in_flight(0)
    # and with nothing waiting in between: only the last two frames are still held by the two slots
    pipe2 = FramePipeline(gpu, W, H, depth=2, to_host=to_host)
    pipe2.begin()
    for t in times:
        pipe2.submit(prm, cam, fx, sky, t)
    pipe2.end()
    torch.cuda.synchronize()
    f = pipe2.last_frame()
    assert np.array_equal(f.numpy() if to_host else f.cpu().numpy(), want[-1])


def test_path_sequence_matches_single_frames_and_sink(gpu, sky_small, tmp_path):
    from relativisticraytracer_b200.parallel import PathSequence
    rrt, sky, prm, fx = _setup(gpu, sky_small)
    n = 7
    path = 0   # "Gargantua Fly-By", src/camera_paths.cpp:33-43
    p = str(tmp_path / "path.rgba")
    seq = PathSequence(gpu, W, H, depth=2)
    l0 = gpu.kernel_launches()
    with rrt.FrameSink(p, W, H) as sink:
        done, launches = seq.render(path, n, prm, fx, sky, fps=24.0, sink=sink)
        assert sink.frames == n
    # launches = the context's own kernel count: 1 per frame from the fused kernel, 3 per pass + 1 from the split pipeline
    assert done == n and launches >= n and launches == gpu.kernel_launches() - l0
    raw = np.fromfile(p, np.uint8).reshape(n, H, W, 4)
    for k in range(1, n + 1):
        t = rrt.path_clock(k, 24.0)                      # recorder clock, src/main.cpp:511-516
        cam, _ = rrt.path_state(path, t)                 # src/main.cpp:176-203
        out = np.zeros((H, W, 4), np.uint8)
        gpu.render_host(prm, cam, fx, sky, t, W, H, out)
        assert np.array_equal(raw[k - 1], out), k
