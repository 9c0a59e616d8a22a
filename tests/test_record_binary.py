"""rrt_record: the native headless recorder (csrc/rrt_record.cpp) -- the reference's path-playback + recording
session (src/main.cpp:171-220, 505-528) written against the C ABI only, no Python in the loop."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "relativisticraytracer_b200", "rrt_record")


def test_usage_and_loud_failure_without_a_device(built):
    assert os.path.exists(BIN)
    assert subprocess.run([BIN], capture_output=True).returncode == 2
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([BIN, "0", "1", "32", "18", "/dev/null"], capture_output=True, text=True)
        assert res.returncode == 1 and "no CUDA device" in res.stderr      # no CPU fallback


@pytest.mark.gpu
def test_recorded_stream_equals_frames_rendered_one_by_one(gpu, sky_small, tmp_path):
    import relativisticraytracer_b200 as rrt
    w, h, n, path = 96, 54, 5, 2                      # "Horizon Skimmer", src/camera_paths.cpp:60-72
    sky_file, out = str(tmp_path / "sky.rgba"), str(tmp_path / "rec.rgba")
    sky_small.tofile(sky_file)
    res = subprocess.run([BIN, str(path), str(n), str(w), str(h), out, "--spin", "0.99", "--sky-raw", sky_file,
                          str(sky_small.shape[1]), str(sky_small.shape[0])], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert f"{n} frames" in res.stdout
    raw = np.fromfile(out, np.uint8).reshape(n, h, w, 4)
    sky = gpu.create_sky(sky_small)
    prm, fx = rrt.default_params(spin_a=0.99), rrt.default_effects()
    for k in range(1, n + 1):
        t = rrt.path_clock(k, 24.0)
        cam, _ = rrt.path_state(path, t)
        want = np.zeros((h, w, 4), np.uint8)
        gpu.render_host(prm, cam, fx, sky, t, w, h, want)
        assert np.array_equal(raw[k - 1], want), k
    sky.close()
    # and through a pipe, like the reference's popen("ffmpeg ...")
    piped = str(tmp_path / "piped.rgba")
    res = subprocess.run([BIN, str(path), "2", str(w), str(h), f"|cat > {piped}", "--spin", "0.99", "--sky-raw", sky_file,
                          str(sky_small.shape[1]), str(sky_small.shape[0])], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert np.array_equal(np.fromfile(piped, np.uint8), raw[:2].ravel())


@pytest.mark.gpu
def test_record_with_a_png_sky_decoded_natively(gpu, sky_small, tmp_path):
    """--sky file.png goes through rrt_sky_load (the native stb-exact decoder + the texture recipe of
    src/main.cpp:246-263): the recorded frame equals a frame rendered from the same pixels handed over as an array."""
    from PIL import Image
    import relativisticraytracer_b200 as rrt
    w, h, path = 96, 54, 0
    png, out = str(tmp_path / "sky.png"), str(tmp_path / "rec.rgba")
    Image.fromarray(sky_small[..., :3]).save(png, "PNG")
    res = subprocess.run([BIN, str(path), "1", str(w), str(h), out, "--spin", "0.99", "--sky", png], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    sky = gpu.create_sky(rrt.load_skybox(png))
    t = rrt.path_clock(1, 24.0)
    cam, _ = rrt.path_state(path, t)
    want = np.zeros((h, w, 4), np.uint8)
    gpu.render_host(rrt.default_params(spin_a=0.99), cam, rrt.default_effects(), sky, t, w, h, want)
    assert np.array_equal(np.fromfile(out, np.uint8).reshape(h, w, 4), want)
    sky.close()
    bad = subprocess.run([BIN, "0", "1", "32", "18", "/dev/null", "--sky", str(tmp_path / "none.jpg")], capture_output=True, text=True)
    assert bad.returncode == 1 and "rrt_sky_load" in bad.stderr
