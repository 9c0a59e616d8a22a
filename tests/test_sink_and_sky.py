"""Host rows either side of the hot path: the frame sink that replaces the
reference's ScreenRecorder (src/main.cpp:29-124) and the skybox file loader (src/main.cpp:237-245).  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest


@pytest.fixture()
def frames():
    rng = np.random.default_rng(7)
    return [rng.integers(0, 256, (6, 10, 4), dtype=np.uint8) for _ in range(3)]


def test_ffmpeg_command_is_the_recorders(built):
    import relativisticraytracer_b200 as rrt
    # src/main.cpp:61-72 with WINDOW_WIDTH x WINDOW_HEIGHT = 1000x700 and RECORDING_FPS 24 (config.h:7-9)
    want = ('ffmpeg -y -f rawvideo -pix_fmt rgba -s 1000x700 -r 24 -i - -vf vflip -c:v libx264 -preset fast '
            '-crf 18 -pix_fmt yuv420p "recording_20260101_000000.mp4"')
    assert rrt.ffmpeg_command(1000, 700, 24, "recording_20260101_000000.mp4") == want


def test_raw_sink_is_the_wire_format(built, frames, tmp_path):
    import relativisticraytracer_b200 as rrt
    p = str(tmp_path / "out.rgba")
    with rrt.FrameSink(p, 10, 6) as s:
        for f in frames:
            s.write(f)
        assert s.frames == 3
    assert open(p, "rb").read() == b"".join(f.tobytes() for f in frames)   # frames back to back, buffer row 0 first


def test_piped_sink_like_popen(built, frames, tmp_path):
    import relativisticraytracer_b200 as rrt
    p = str(tmp_path / "piped.rgba")
    with rrt.FrameSink(f"|cat > {p}", 10, 6) as s:
        s.write(frames[0])
    assert open(p, "rb").read() == frames[0].tobytes()
    # a command that cannot be started / fails reports an I/O error on close, it does not pass silently
    s = rrt.FrameSink("|exit 3", 10, 6)
    with pytest.raises(rrt.RrtError):
        s.write(frames[0])
        s.close()


def test_y4m_sink_flips_rows_and_converts(built, frames, tmp_path):
    import relativisticraytracer_b200 as rrt
    p = str(tmp_path / "out.y4m")
    w, h = 10, 6
    with rrt.FrameSink(p, w, h, fps=24, fmt=rrt.SINK_Y4M) as s:
        for f in frames[:2]:
            s.write(f)
    raw = open(p, "rb").read()
    head = b"YUV4MPEG2 W10 H6 F24:1 Ip A1:1 C420jpeg\n"
    assert raw.startswith(head)
    fsz = w * h + 2 * (w // 2) * (h // 2)
    assert len(raw) == len(head) + 2 * (6 + fsz)
    body = raw[len(head):]
    for i, f in enumerate(frames[:2]):
        chunk = body[i * (6 + fsz):(i + 1) * (6 + fsz)]
        assert chunk[:6] == b"FRAME\n"
        y = np.frombuffer(chunk[6:6 + w * h], np.uint8).reshape(h, w)
        fl = f[::-1].astype(np.int64)                                   # -vf vflip
        want_y = ((66 * fl[..., 0] + 129 * fl[..., 1] + 25 * fl[..., 2] + 128) >> 8) + 16
        assert np.array_equal(y, want_y.astype(np.uint8))
        u = np.frombuffer(chunk[6 + w * h:6 + w * h + (w // 2) * (h // 2)], np.uint8).reshape(h // 2, w // 2)
        blk = (fl[..., :3].reshape(h // 2, 2, w // 2, 2, 3).sum(axis=(1, 3)) + 2) >> 2
        want_u = ((-38 * blk[..., 0] - 74 * blk[..., 1] + 112 * blk[..., 2] + 128) >> 8) + 128
        assert np.array_equal(u, want_u.astype(np.uint8))


def test_sink_bad_arguments(built, tmp_path):
    from relativisticraytracer_b200 import _capi
    lib = _capi.load()
    h = C.c_void_p()
    assert lib.rrt_sink_open(None, 0, 4, 4, 24, C.byref(h)) == _capi.ERR_BAD_ARG
    assert lib.rrt_sink_open(b"x", 7, 4, 4, 24, C.byref(h)) == _capi.ERR_BAD_ARG
    assert lib.rrt_sink_open(b"x", 0, 0, 4, 24, C.byref(h)) == _capi.ERR_BAD_ARG
    assert lib.rrt_sink_open(str(tmp_path / "no" / "such" / "dir" / "f").encode(), 0, 4, 4, 24, C.byref(h)) == _capi.ERR_IO
    assert lib.rrt_sink_write(None, None) == _capi.ERR_BAD_ARG
    assert lib.rrt_sink_close(None) == _capi.ERR_BAD_ARG
    buf = C.create_string_buffer(8)
    assert lib.rrt_sink_ffmpeg_command(10, 10, 24, b"o.mp4", buf, 8) == _capi.ERR_BAD_ARG   # buffer too small


def test_load_skybox_formats(built, tmp_path):
    import relativisticraytracer_b200 as rrt
    from PIL import Image
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (8, 16, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(tmp_path / "s.png")
    got = rrt.load_skybox(str(tmp_path / "s.png"))
    assert got.shape == (8, 16, 4) and got.dtype == np.uint8
    assert np.array_equal(got[..., :3], rgb) and (got[..., 3] == 255).all()      # stbi req_comp=4: alpha = 255
    rgba = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    Image.fromarray(rgba).save(tmp_path / "a.png")
    assert np.array_equal(rrt.load_skybox(str(tmp_path / "a.png")), rgba)
    grey = rng.integers(0, 256, (8, 16), dtype=np.uint8)
    Image.fromarray(grey).save(tmp_path / "g.png")
    g = rrt.load_skybox(str(tmp_path / "g.png"))
    assert all(np.array_equal(g[..., c], grey) for c in range(3)) and (g[..., 3] == 255).all()
    np.save(tmp_path / "s.npy", rgba)
    assert np.array_equal(rrt.load_skybox(str(tmp_path / "s.npy")), rgba)
    rgba.tofile(tmp_path / "s.rgba")
    assert np.array_equal(rrt.load_skybox(str(tmp_path / "s.rgba"), 16, 8), rgba)
    with pytest.raises(ValueError):
        rrt.load_skybox(str(tmp_path / "s.rgba"))
    with pytest.raises(ValueError):
        rrt.load_skybox(str(tmp_path / "s.rgba"), 15, 8)
