"""Frame sink: the step after the hot path (reference ScreenRecorder, src/main.cpp:29-124), headless.

``FrameSink(target, w, h, fps, fmt)`` writes host frames (``[h, w, 4]`` uint8, exactly what ``render_host`` /
``launch_raymarch`` produce) either as the recorder's raw ``rgba`` wire format -- to a file, or to a command
with ``target="|ffmpeg ..."`` like the reference's ``popen`` -- or as a YUV4MPEG2 file.  ``ffmpeg_command``
returns the reference's own command line (src/main.cpp:61-72)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import RrtError, SINK_RGBA, SINK_Y4M


def ffmpeg_command(w: int, h: int, fps: int = 24, out_name: str = "recording.mp4") -> str:
    buf = C.create_string_buffer(1024)
    n = _capi.load().rrt_sink_ffmpeg_command(int(w), int(h), int(fps), out_name.encode(), buf, len(buf))
    if n < 0:
        raise RrtError(n, "rrt_sink_ffmpeg_command")
    return buf.value.decode()


class FrameSink:
    def __init__(self, target: str, w: int, h: int, fps: int = 24, fmt: int = SINK_RGBA):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.w, self.h = int(w), int(h)
        rc = self._lib.rrt_sink_open(target.encode(), int(fmt), self.w, self.h, int(fps), C.byref(self._h))
        if rc != 0:
            raise RrtError(rc, f"rrt_sink_open({target!r})")

    def write(self, frame) -> None:
        """frame: numpy array or CPU torch tensor, contiguous, h*w*4 bytes."""
        if hasattr(frame, "data_ptr"):
            assert not frame.is_cuda and frame.is_contiguous() and frame.numel() * frame.element_size() == self.w * self.h * 4
            ptr = frame.data_ptr()
        else:
            frame = np.ascontiguousarray(frame, dtype=np.uint8)
            assert frame.nbytes == self.w * self.h * 4
            ptr = frame.ctypes.data
        rc = self._lib.rrt_sink_write(self._h, C.c_void_p(ptr))
        if rc != 0:
            raise RrtError(rc, "rrt_sink_write")

    @property
    def frames(self) -> int:
        return int(self._lib.rrt_sink_frames(self._h))

    def close(self) -> None:
        if self._h:
            h, self._h = self._h, C.c_void_p()
            rc = self._lib.rrt_sink_close(h)
            if rc != 0:
                raise RrtError(rc, "rrt_sink_close")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
