"""Host-side mirror of the reference's render interface over the C ABI.

The reference exposes one call, ``launch_raymarch(d_out, w, h, time, cam, skyboxTex, effects)``
(include/raymarcher.h:19), fed by ``CameraController::getCUDAStateFrom`` (src/main.cpp:141-167) and
``loadSkybox`` (src/main.cpp:237-266).  This module keeps those names and argument meanings.  PyTorch is
used for what it is good at here -- device buffers, streams and (in ``parallel.py``) ``torch.distributed``;
every pixel is computed by the hand-written sm_100a kernels in ``csrc/`` through ``librrt_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _capi
from ._capi import (Band, Camera, Counters, Effects, Params, Planes, RrtError, OUT_FRAME, OUT_PACKED,
                    default_effects, default_params, effects_off)

CameraState = Camera      # reference name (include/raymarcher.h:11)
CameraEffects = Effects   # reference name (camera_settings.h:4)


def camera_state_from(pos, yaw_deg: float, pitch_deg: float) -> Camera:
    """CameraController::getCUDAStateFrom (src/main.cpp:141-167); angles in degrees.  Host only."""
    cam = Camera()
    arr = (C.c_float * 3)(*[float(x) for x in pos])
    _capi.load().rrt_camera_from(C.byref(arr), float(yaw_deg), float(pitch_deg), C.byref(cam))
    return cam


def path_names():
    lib = _capi.load()
    return [lib.rrt_path_name(i).decode() for i in range(lib.rrt_path_count())]


def path_state(path_index: int, t: float):
    """PathController::getInterpolatedState (src/main.cpp:176-203) -> (Camera, [x,y,z,yaw,pitch])."""
    cam = Camera()
    pyp = np.zeros(5, np.float32)
    rc = _capi.load().rrt_path_state(int(path_index), float(t), C.byref(cam), pyp.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RrtError(rc, f"rrt_path_state({path_index}, {t})")
    return cam, pyp


def path_clock(frame: int, fps: float = 24.0) -> float:
    """pathTime after `frame` float accumulations of 1/fps (recording clock, src/main.cpp:511-516)."""
    return float(_capi.load().rrt_path_clock(int(frame), float(fps)))


def path_duration(path_index: int) -> float:
    return float(_capi.load().rrt_path_duration(int(path_index)))


class Sky:
    """A skybox texture object (rrt_sky) -- the device half of the reference's loadSkybox."""

    def __init__(self, renderer: "Renderer", rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        if rgba.ndim != 3 or rgba.shape[2] != 4:
            raise ValueError("sky must be [h, w, 4] uint8 (RGBA8, rows top-down as stbi_load returns them)")
        self._r = renderer
        self._h = C.c_void_p()
        self.height, self.width = int(rgba.shape[0]), int(rgba.shape[1])
        renderer._check(renderer._lib.rrt_sky_create(renderer._ctx, rgba.ctypes.data_as(C.c_void_p), self.width,
                                                     self.height, C.byref(self._h)))
        self.texture = int(renderer._lib.rrt_sky_texture(self._h))   # cudaTextureObject_t

    def close(self):
        if self._h:
            self._r._lib.rrt_sky_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PLANE_SPECS = {"hdr": (4, torch.float32), "dir": (4, torch.float32), "emis": (4, torch.float32),
               "pos": (4, torch.float32), "vel": (4, torch.float32), "cls": (0, torch.uint8), "steps": (0, torch.int32)}


class PeerFrame:
    """A device frame on the encoding GPU that other processes' render kernels store into directly (C ABI
    rrt_peer_frame_*; CUDA IPC).  The owner (rank 0) creates it and passes ``handle`` (64 bytes) to the other ranks,
    which ``open`` it; the owner reads the assembled frame out with ``read_into`` (pinned host or device tensor)."""

    def __init__(self, renderer: "Renderer", h: int, w: int, handle: Optional[bytes] = None):
        self.r, self.h, self.w, self.nbytes = renderer, h, w, h * w * 4
        self.owner = handle is None
        ptr = C.c_void_p()
        if self.owner:
            buf = (C.c_uint8 * 64)()
            renderer._check(renderer._lib.rrt_peer_frame_create(renderer._ctx, C.c_size_t(self.nbytes), C.byref(ptr), buf))
            self.handle = bytes(buf)
        else:
            assert len(handle) == 64
            self.handle = bytes(handle)
            buf = (C.c_uint8 * 64).from_buffer_copy(self.handle)
            renderer._check(renderer._lib.rrt_peer_frame_open(renderer._ctx, buf, C.byref(ptr)))
        self.ptr = int(ptr.value)

    def read_into(self, dst: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Stream-ordered copy of the frame into `dst` (pinned host or device uint8 tensor of the frame's size)."""
        assert dst.dtype == torch.uint8 and dst.is_contiguous() and dst.numel() == self.nbytes
        st = stream if stream is not None else torch.cuda.current_stream(self.r.device)
        self.r._check(self.r._lib.rrt_peer_frame_read(self.r._ctx, C.c_void_p(self.ptr), C.c_size_t(self.nbytes),
                                                      C.c_void_p(dst.data_ptr()), C.c_void_p(st.cuda_stream)))

    def close(self):
        if self.ptr:
            self.r._lib.rrt_peer_frame_close(self.r._ctx, C.c_void_p(self.ptr), 1 if self.owner else 0)
            self.ptr = 0


class Renderer:
    """One rrt_context on one B200.  Thread-compatible; launches are asynchronous on the given stream."""

    def __init__(self, device: int | None = None):
        self._lib = _capi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("relativisticraytracer_b200 needs a CUDA device (no CPU fallback exists)")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self._ctx = C.c_void_p()
        rc = self._lib.rrt_context_create(self.device_index, C.byref(self._ctx))
        if rc != 0:
            raise RrtError(rc, (self._lib.rrt_last_error(None) or b"").decode())

    # ---- plumbing -------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise RrtError(rc, (self._lib.rrt_last_error(self._ctx) or b"").decode())

    def close(self):
        if self._ctx:
            self._lib.rrt_context_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create_sky(self, rgba: np.ndarray) -> Sky:
        return Sky(self, rgba)

    @staticmethod
    def band_rows(band: Optional[Band], h: int) -> int:
        return int(_capi.load().rrt_band_rows(C.byref(band) if band is not None else None, int(h)))

    def alloc_planes(self, w: int, h: int, names=("hdr", "dir", "emis", "pos", "vel", "cls", "steps")):
        out = {}
        for n in names:
            ch, dt = PLANE_SPECS[n]
            shape = (h, w, ch) if ch else (h, w)
            out[n] = torch.zeros(shape, dtype=dt, device=self.device)
        return out

    # ---- the hot path -----------------------------------------------------------------------------
    def render(self, prm: Params, cam: Camera, fx: Effects, sky: Sky | int, time: float, w: int, h: int, *,
               band: Optional[Band] = None, out: Optional[torch.Tensor] = None, layout: int = OUT_FRAME,
               planes: Optional[dict] = None, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """rrt_render: one frame (or one band of it) into a device uchar4 tensor; asynchronous."""
        tex = sky.texture if isinstance(sky, Sky) else int(sky)
        rows = h if layout == OUT_FRAME else self.band_rows(band, h)
        if isinstance(out, PeerFrame):      # a frame that lives on the encoding GPU (possibly another process's)
            assert layout == OUT_FRAME and out.nbytes >= h * w * 4
            out_ptr = out.ptr
        else:
            if out is None:
                out = torch.empty((rows, w, 4), dtype=torch.uint8, device=self.device)
            assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.numel() >= rows * w * 4
            out_ptr = out.data_ptr()
        pl = None
        if planes:
            pl = Planes()
            for n, t in planes.items():
                ch, dt = PLANE_SPECS[n]
                assert t.is_cuda and t.dtype == dt and t.is_contiguous() and t.numel() == h * w * max(ch, 1), n
                setattr(pl, n, t.data_ptr())
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._check(self._lib.rrt_render(self._ctx, C.byref(prm), C.byref(cam), C.byref(fx), C.c_uint64(tex),
                                         float(time), int(w), int(h), C.byref(band) if band is not None else None,
                                         C.c_void_p(out_ptr), int(layout),
                                         C.byref(pl) if pl is not None else None, C.c_void_p(st.cuda_stream)))
        return out

    def render_host(self, prm: Params, cam: Camera, fx: Effects, sky: Sky | int, time: float, w: int, h: int,
                    host_out) -> None:
        """rrt_render_host: end-to-end call, HOST destination (numpy array or pinned torch tensor)."""
        tex = sky.texture if isinstance(sky, Sky) else int(sky)
        if isinstance(host_out, torch.Tensor):
            assert not host_out.is_cuda and host_out.is_contiguous() and host_out.numel() * host_out.element_size() >= w * h * 4
            ptr = host_out.data_ptr()
        else:
            assert host_out.flags["C_CONTIGUOUS"] and host_out.nbytes >= w * h * 4
            ptr = host_out.ctypes.data
        self._check(self._lib.rrt_render_host(self._ctx, C.byref(prm), C.byref(cam), C.byref(fx), C.c_uint64(tex),
                                              float(time), int(w), int(h), C.c_void_p(ptr)))

    def render_host_async(self, prm: Params, cam: Camera, fx: Effects, sky: Sky | int, time: float, w: int, h: int,
                          host_out: torch.Tensor, slot: int = 0, stream: Optional[torch.cuda.Stream] = None) -> None:
        """rrt_render_host_async: trace + device->host copy enqueued on `stream`, no synchronisation.
        ``host_out`` must be a pinned CPU tensor; ``slot`` < HOST_SLOTS selects the context's device frame."""
        tex = sky.texture if isinstance(sky, Sky) else int(sky)
        assert (not host_out.is_cuda) and host_out.is_pinned() and host_out.is_contiguous()
        assert host_out.numel() * host_out.element_size() >= w * h * 4
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._check(self._lib.rrt_render_host_async(self._ctx, C.byref(prm), C.byref(cam), C.byref(fx), C.c_uint64(tex),
                                                    float(time), int(w), int(h), C.c_void_p(host_out.data_ptr()),
                                                    int(slot), C.c_void_p(st.cuda_stream)))

    def assemble_bands(self, packed: torch.Tensor, rows_per_rank: int, w: int, h: int, nranks: int, group: int,
                       frame: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        if frame is None:
            frame = torch.empty((h, w, 4), dtype=torch.uint8, device=self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._check(self._lib.rrt_assemble_bands(self._ctx, C.c_void_p(packed.data_ptr()), int(rows_per_rank), int(w),
                                                 int(h), int(nranks), int(group), C.c_void_p(frame.data_ptr()),
                                                 C.c_void_p(st.cuda_stream)))
        return frame

    def read_counters(self, reset: bool = True) -> dict:
        c = Counters()
        self._check(self._lib.rrt_read_counters(self._ctx, C.byref(c), 1 if reset else 0))
        return c.as_dict()

    def fp32_peak(self, iters: int = 4096):
        tf, ms = C.c_double(), C.c_double()
        self._check(self._lib.rrt_fp32_peak_probe(self._ctx, int(iters), C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def set_frames_in_flight(self, n: int) -> None:
        """rrt_set_frames_in_flight: each launch takes 1/n of the resident-CTA slots (n concurrent frames)."""
        self._check(self._lib.rrt_set_frames_in_flight(self._ctx, int(n)))

    def set_pipeline(self, mode) -> None:
        """rrt_set_pipeline: "auto" (split whenever a medium is on), "fused" (one kernel) or "split"
        (trace / media / fold kernels over the sample pool); same frames either way."""
        modes = {"auto": _capi.PIPELINE_AUTO, "fused": _capi.PIPELINE_FUSED, "split": _capi.PIPELINE_SPLIT}
        self._check(self._lib.rrt_set_pipeline(self._ctx, int(modes.get(mode, mode))))

    def set_sample_pool(self, max_bytes_per_stream: int = 0, max_passes: int = 0) -> None:
        """rrt_set_sample_pool: size limit of one sample pool and the passes a frame may be cut into (0 = unchanged)."""
        self._check(self._lib.rrt_set_sample_pool(self._ctx, C.c_size_t(int(max_bytes_per_stream)), int(max_passes)))

    def kernel_launches(self) -> int:
        """rrt_kernel_launches: kernels launched so far by this context's render / assemble calls."""
        return int(self._lib.rrt_kernel_launches(self._ctx))

    def split_stats(self) -> dict:
        """rrt_split_stats: bookkeeping of the last frame the split pipeline completed (synchronises)."""
        out = (C.c_uint32 * 8)()
        self._check(self._lib.rrt_split_stats(self._ctx, out))
        return {"passes_worked": int(out[0]), "tiles_swept": int(out[1]), "tiles_split": int(out[2]),
                "passes_enqueued": int(out[3]), "tiles": int(out[4]), "pool_kislots": int(out[5])}

    def tile_log(self, log: Optional[torch.Tensor]) -> None:
        """rrt_debug_tile_log: per-tile (start ns, end ns, row<<32|col, sm<<32|max steps) into `log` ([n, 4] int64/uint64
        device tensor), or None to switch it off."""
        if log is None:
            self._check(self._lib.rrt_debug_tile_log(self._ctx, None, 0))
        else:
            assert log.is_cuda and log.is_contiguous() and log.element_size() == 8 and log.shape[-1] == 4
            self._check(self._lib.rrt_debug_tile_log(self._ctx, C.c_void_p(log.data_ptr()), C.c_size_t(log.shape[0])))

    def set_probe_contract(self, fmad: bool) -> None:
        """Rounding contract of hash31 / noise3d / fbm (the probes without a parameter block)."""
        self._check(self._lib.rrt_set_probe_contract(self._ctx, 1 if fmad else 0))

    def exact_math_selftest(self, seed: int = 1, n: int = 1 << 30):
        """(div mismatches, sqrt mismatches) of the loop's branch-free div/sqrt vs the IEEE intrinsics."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.rrt_exact_math_selftest(self._ctx, C.c_uint64(seed), C.c_uint64(n), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def exact_pow_selftest(self, seed: int = 1, n: int = 1 << 30):
        """rrt_exact_pow_selftest: (mismatches with the media code's exponents, mismatches with random exponents)."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.rrt_exact_pow_selftest(self._ctx, C.c_uint64(seed), C.c_uint64(n), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    # ---- function-level probes (host numpy in / out) --------------------------------------------------
    @staticmethod
    def _f32(a):
        return np.ascontiguousarray(np.asarray(a, dtype=np.float32))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def geodesic_acc(self, prm, q, v):
        q, v = self._f32(q), self._f32(v)
        out = np.empty_like(q)
        self._check(self._lib.rrt_geodesic_acc_batch(self._ctx, C.byref(prm), len(q), self._p(q), self._p(v), self._p(out)))
        return out

    def _step(self, fn, prm, p, v, h):
        p, v = self._f32(p).copy(), self._f32(v).copy()
        h = self._f32(np.broadcast_to(np.asarray(h, np.float32), (len(p),)))
        self._check(fn(self._ctx, C.byref(prm), len(p), self._p(p), self._p(v), self._p(h)))
        return p, v

    def rk4_step(self, prm, p, v, h):
        return self._step(self._lib.rrt_rk4_step_batch, prm, p, v, h)

    def euler_step(self, prm, p, v, h):
        return self._step(self._lib.rrt_euler_step_batch, prm, p, v, h)

    def redshift(self, prm, q, v):
        q, v = self._f32(q), self._f32(v)
        out = np.empty(len(q), np.float32)
        self._check(self._lib.rrt_redshift_batch(self._ctx, C.byref(prm), len(q), self._p(q), self._p(v), self._p(out)))
        return out

    def hash31(self, p):
        p = self._f32(p)
        out = np.empty(len(p), np.float32)
        self._check(self._lib.rrt_hash31_batch(self._ctx, len(p), self._p(p), self._p(out)))
        return out

    def noise3d(self, p):
        p = self._f32(p)
        out = np.empty(len(p), np.float32)
        self._check(self._lib.rrt_noise3d_batch(self._ctx, len(p), self._p(p), self._p(out)))
        return out

    def fbm(self, p, octaves: int):
        p = self._f32(p)
        out = np.empty(len(p), np.float32)
        self._check(self._lib.rrt_fbm_batch(self._ctx, len(p), self._p(p), int(octaves), self._p(out)))
        return out

    def disk_temperature(self, prm, r):
        r = self._f32(r)
        out = np.empty_like(r)
        self._check(self._lib.rrt_disk_temperature_batch(self._ctx, C.byref(prm), len(r), self._p(r), self._p(out)))
        return out

    def disk_density(self, prm, q, time):
        q = self._f32(q)
        out = np.empty(len(q), np.float32)
        self._check(self._lib.rrt_disk_density_batch(self._ctx, C.byref(prm), len(q), self._p(q), float(time), self._p(out)))
        return out

    def dust_density(self, prm, q, time):
        q = self._f32(q)
        out = np.empty(len(q), np.float32)
        self._check(self._lib.rrt_dust_density_batch(self._ctx, C.byref(prm), len(q), self._p(q), float(time), self._p(out)))
        return out

    def sky_sample(self, sky: Sky | int, tx, ty):
        tex = sky.texture if isinstance(sky, Sky) else int(sky)
        tx, ty = self._f32(tx), self._f32(ty)
        out = np.empty((len(tx), 4), np.float32)
        self._check(self._lib.rrt_sky_sample_batch(self._ctx, C.c_uint64(tex), len(tx), self._p(tx), self._p(ty), self._p(out)))
        return out


# ---- the reference's own entry point, same name and argument order (include/raymarcher.h:19) ------------
_default: dict[int, Renderer] = {}
_launch_params: Optional[Params] = None


def set_launch_params(prm: Optional[Params]) -> None:
    """Parameter block used by launch_raymarch (the reference compiles these in; default = config.h)."""
    global _launch_params
    _launch_params = prm


def launch_raymarch(d_out: torch.Tensor, w: int, h: int, time: float, cam: Camera, skyboxTex, effects: Effects) -> None:
    """Drop-in for the reference launcher: renders into the caller-owned device buffer ``d_out``
    (uchar4 per pixel, pixel (x,y) at [(h-1-y)*w + x]) on the current stream, asynchronously."""
    dev = d_out.device.index if d_out.device.index is not None else torch.cuda.current_device()
    r = _default.get(dev)
    if r is None:
        r = _default[dev] = Renderer(dev)
    prm = _launch_params if _launch_params is not None else default_params()
    r.render(prm, cam, effects, skyboxTex, time, w, h, out=d_out, layout=OUT_FRAME)
