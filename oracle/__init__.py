"""ctypes front-end for the CPU checkers (oracle/oracle_abi.h).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs -- never from the product package
``relativisticraytracer_b200`` (tests/test_layout.py enforces that).

Two libraries implement the same ABI:

* ``Oracle("port")``      -> oracle/librrt_oracle.so   (plain-C restatement, built by ``make port``)
* ``Oracle("reference")`` -> oracle/_ref/libref_host.so (the reference's own headers compiled for the
  host by ``make ref``; only buildable where /root/reference exists, the built .so travels)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TREE = os.environ.get("RRT_REFERENCE_TREE", "/root/reference")

FLAG_DISK = 1
FLAG_DUST = 2
CLS_CAPTURED, CLS_DISK_HIT, CLS_ESCAPED = 0, 1, 2
CLS_MASK = 3
CLSF_EXHAUSTED = 4
CLSF_TOUCHED = 8


class Params(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "spin_a", "event_horizon", "isco_radius", "disk_out", "disk_h", "disk_luminosity", "disk_opacity",
        "exposure", "cloud_h", "cloud_out", "cloud_opacity", "cloud_luminosity", "step_size", "disk_temp_ref")] + [
        ("max_steps", C.c_int32), ("flags", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("forward", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3)]


class Effects(C.Structure):
    _fields_ = [("use_bloom", C.c_int32), ("bloom_threshold", C.c_float), ("bloom_intensity", C.c_float),
                ("use_vignette", C.c_int32), ("vignette_intensity", C.c_float),
                ("use_ca", C.c_int32), ("ca_amount", C.c_float),
                ("use_lens", C.c_int32), ("distortion_amount", C.c_float)]


class Planes(C.Structure):
    _fields_ = [("hdr", C.c_void_p), ("dir", C.c_void_p), ("emis", C.c_void_p), ("pos", C.c_void_p),
                ("vel", C.c_void_p), ("cls", C.c_void_p), ("steps", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rk4_steps", "disk_evals", "dust_evals", "dense_samples",
                                          "n_captured", "n_escaped", "n_exhausted", "n_touched")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


@dataclass
class Frame:
    rgba: np.ndarray   # [h, w, 4] uint8, row-flipped exactly like the reference store
    hdr: np.ndarray    # [h, w, 4] float32 (rgb, transmittance)
    dir: np.ndarray    # [h, w, 4]
    emis: np.ndarray   # [h, w, 4]
    pos: np.ndarray    # [h, w, 4]
    vel: np.ndarray    # [h, w, 4]
    cls: np.ndarray    # [h, w] uint8
    steps: np.ndarray  # [h, w] int32
    counters: dict


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def build(kind: str = "port", quiet: bool = True) -> str:
    """Compile the requested checker with oracle/Makefile; returns the library path."""
    target = {"port": "port", "reference": "ref", "reference_fast": "ref_fast", "reference_cuda": "ref_cuda"}[kind]
    path = {"port": os.path.join(HERE, "librrt_oracle.so"),
            "reference": os.path.join(HERE, "_ref", "libref_host.so"),
            "reference_fast": os.path.join(HERE, "_ref", "libref_host_fast.so"),
            "reference_cuda": os.path.join(HERE, "_ref", "libref_cuda.so")}[kind]
    if kind != "port" and not os.path.isdir(os.path.join(REF_TREE, "include")):
        if os.path.exists(path):
            return path  # prebuilt binary travelled with the snapshot
        raise FileNotFoundError(f"{path} not built and reference tree {REF_TREE} absent")
    res = subprocess.run(["make", "-C", HERE, target, f"REF={REF_TREE}"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"make {target} failed:\n{res.stdout}\n{res.stderr}")
    if not quiet:
        print(res.stdout)
    return path


def _cpu_has(*flags) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    have = set(line.split(":", 1)[1].split())
                    return all(x in have for x in flags)
    except OSError:
        pass
    return False


def available(kind: str) -> bool:
    """"reference_fast" = the reference headers built -O3 -mavx2 -mfma with contraction (timing baseline only, never
    a parity witness); it is reported available only on a host CPU that has AVX2 and FMA."""
    path = {"port": os.path.join(HERE, "librrt_oracle.so"),
            "reference": os.path.join(HERE, "_ref", "libref_host.so"),
            "reference_fast": os.path.join(HERE, "_ref", "libref_host_fast.so")}[kind]
    if kind == "reference_fast" and not _cpu_has("avx2", "fma"):
        return False
    return os.path.exists(path)


class Oracle:
    """One of the two CPU checkers behind the oracle ABI."""

    def __init__(self, kind: str = "port", auto_build: bool = True):
        assert kind in ("port", "reference", "reference_fast")
        self.kind = kind
        self.prefix = "ora_" if kind == "port" else "ref_"
        path = {"port": os.path.join(HERE, "librrt_oracle.so"), "reference": os.path.join(HERE, "_ref", "libref_host.so"),
                "reference_fast": os.path.join(HERE, "_ref", "libref_host_fast.so")}[kind]
        if auto_build:
            try:
                path = build(kind)
            except FileNotFoundError:
                raise
        self.path = path
        self.lib = C.CDLL(path)
        self._sig()

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _sig(self):
        P = C.POINTER
        vp = C.c_void_p
        self._fn("default_params").argtypes = [P(Params)]
        self._fn("default_effects").argtypes = [P(Effects)]
        self._fn("num_threads").restype = C.c_int
        self._fn("set_num_threads").argtypes = [C.c_int]
        self._fn("camera_from").argtypes = [P(C.c_float * 3), C.c_float, C.c_float, P(Camera)]
        self._fn("path_state").argtypes = [C.c_int, C.c_float, P(Camera), vp]
        self._fn("path_state").restype = C.c_int
        self._fn("render").argtypes = [P(Params), P(Camera), P(Effects), vp, C.c_int, C.c_int, C.c_float,
                                       C.c_int, C.c_int, C.c_int, C.c_int, vp, P(Planes), P(Counters)]
        self._fn("render").restype = C.c_int
        self._fn("geodesic_acc").argtypes = [P(Params), C.c_int, vp, vp, vp]
        self._fn("rk4_step").argtypes = [P(Params), C.c_int, vp, vp, vp]
        self._fn("euler_step").argtypes = [P(Params), C.c_int, vp, vp, vp]
        self._fn("redshift").argtypes = [P(Params), C.c_int, vp, vp, vp]
        self._fn("hash31").argtypes = [C.c_int, vp, vp]
        self._fn("noise3d").argtypes = [C.c_int, vp, vp]
        self._fn("fbm").argtypes = [C.c_int, vp, C.c_int, vp]
        self._fn("disk_temperature").argtypes = [P(Params), C.c_int, vp, vp]
        self._fn("disk_density").argtypes = [P(Params), C.c_int, vp, C.c_float, vp]
        self._fn("dust_density").argtypes = [P(Params), C.c_int, vp, C.c_float, vp]
        self._fn("tex2d").argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        if self.kind == "port":
            self.lib.ora_set_probe_contract.argtypes = [C.c_int]

    def set_probe_contract(self, fmad: bool) -> None:
        """Contract of hash31 / noise3d / fbm (the probes without a parameter block); port only."""
        assert self.kind == "port"
        self.lib.ora_set_probe_contract(1 if fmad else 0)

    # ---- parameter helpers -------------------------------------------------------------
    def default_params(self, **over) -> Params:
        p = Params()
        self._fn("default_params")(C.byref(p))
        for k, v in over.items():
            setattr(p, k, v)
        return p

    def default_effects(self, **over) -> Effects:
        e = Effects()
        self._fn("default_effects")(C.byref(e))
        for k, v in over.items():
            setattr(e, k, v)
        return e

    def effects_off(self) -> Effects:
        return self.default_effects(use_bloom=0, use_vignette=0, use_ca=0, use_lens=0)

    def num_threads(self) -> int:
        return int(self._fn("num_threads")())

    def set_num_threads(self, n: int) -> None:
        """OpenMP team of the next render; launchers such as torchrun export OMP_NUM_THREADS=1."""
        self._fn("set_num_threads")(int(n))

    def camera_from(self, pos, yaw_deg, pitch_deg) -> Camera:
        cam = Camera()
        arr = (C.c_float * 3)(*[float(x) for x in pos])
        self._fn("camera_from")(C.byref(arr), float(yaw_deg), float(pitch_deg), C.byref(cam))
        return cam

    def path_state(self, path_index: int, t: float):
        cam = Camera()
        pyp = np.zeros(5, np.float32)
        rc = self._fn("path_state")(int(path_index), float(t), C.byref(cam), _ptr(pyp))
        if rc != 0:
            raise ValueError(f"path_state rc={rc}")
        return cam, pyp

    # ---- whole frame ------------------------------------------------------------------
    def render(self, prm: Params, cam: Camera, fx: Effects, sky: np.ndarray, time: float, w: int, h: int,
               y0: int = 0, y1: int | None = None) -> Frame:
        y1 = h if y1 is None else y1
        sky = np.ascontiguousarray(sky, dtype=np.uint8)
        assert sky.ndim == 3 and sky.shape[2] == 4
        f = Frame(rgba=np.zeros((h, w, 4), np.uint8), hdr=np.zeros((h, w, 4), np.float32),
                  dir=np.zeros((h, w, 4), np.float32), emis=np.zeros((h, w, 4), np.float32),
                  pos=np.zeros((h, w, 4), np.float32), vel=np.zeros((h, w, 4), np.float32),
                  cls=np.zeros((h, w), np.uint8), steps=np.zeros((h, w), np.int32), counters={})
        pl = Planes(_ptr(f.hdr), _ptr(f.dir), _ptr(f.emis), _ptr(f.pos), _ptr(f.vel), _ptr(f.cls), _ptr(f.steps))
        cnt = Counters()
        rc = self._fn("render")(C.byref(prm), C.byref(cam), C.byref(fx), _ptr(sky), sky.shape[1], sky.shape[0],
                                float(time), w, h, y0, y1, _ptr(f.rgba), C.byref(pl), C.byref(cnt))
        if rc != 0:
            raise ValueError(f"{self.prefix}render rc={rc}")
        f.counters = cnt.as_dict()
        return f

    # ---- function-level probes --------------------------------------------------------
    def geodesic_acc(self, prm, q, v):
        q, v = _f32(q), _f32(v)
        out = np.empty_like(q)
        self._fn("geodesic_acc")(C.byref(prm), len(q), _ptr(q), _ptr(v), _ptr(out))
        return out

    def _step(self, name, prm, p, v, h):
        p, v = _f32(p).copy(), _f32(v).copy()
        h = _f32(np.broadcast_to(np.asarray(h, np.float32), (len(p),)))
        self._fn(name)(C.byref(prm), len(p), _ptr(p), _ptr(v), _ptr(h))
        return p, v

    def rk4_step(self, prm, p, v, h):
        return self._step("rk4_step", prm, p, v, h)

    def euler_step(self, prm, p, v, h):
        return self._step("euler_step", prm, p, v, h)

    def redshift(self, prm, q, v):
        q, v = _f32(q), _f32(v)
        out = np.empty(len(q), np.float32)
        self._fn("redshift")(C.byref(prm), len(q), _ptr(q), _ptr(v), _ptr(out))
        return out

    def _scalar_of_p(self, name, p, *extra):
        p = _f32(p)
        out = np.empty(len(p), np.float32)
        self._fn(name)(len(p), _ptr(p), *extra, _ptr(out))
        return out

    def hash31(self, p):
        return self._scalar_of_p("hash31", p)

    def noise3d(self, p):
        return self._scalar_of_p("noise3d", p)

    def fbm(self, p, octaves):
        return self._scalar_of_p("fbm", p, int(octaves))

    def disk_temperature(self, prm, r):
        r = _f32(r)
        out = np.empty_like(r)
        self._fn("disk_temperature")(C.byref(prm), len(r), _ptr(r), _ptr(out))
        return out

    def disk_density(self, prm, q, time):
        q = _f32(q)
        out = np.empty(len(q), np.float32)
        self._fn("disk_density")(C.byref(prm), len(q), _ptr(q), float(time), _ptr(out))
        return out

    def dust_density(self, prm, q, time):
        q = _f32(q)
        out = np.empty(len(q), np.float32)
        self._fn("dust_density")(C.byref(prm), len(q), _ptr(q), float(time), _ptr(out))
        return out

    def tex2d(self, sky, tx, ty):
        sky = np.ascontiguousarray(sky, dtype=np.uint8)
        tx, ty = _f32(tx), _f32(ty)
        out = np.empty((len(tx), 4), np.float32)
        self._fn("tex2d")(_ptr(sky), sky.shape[1], sky.shape[0], len(tx), _ptr(tx), _ptr(ty), _ptr(out))
        return out


class RefCuda:
    """The reference's own CUDA raymarcher (src/raymarcher.cu, unmodified, nvcc defaults + sm_100a) behind a
    headless harness (oracle/ref_cuda_harness.cu).  Needs a GPU.  spin must be 0.0 or 0.99 (compile-time there)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_cuda.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.refcuda_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.refcuda_render.restype = C.c_int

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libref_cuda.so"))

    def render(self, spin: float, cam, fx, sky: np.ndarray, time: float, w: int, h: int, reps: int = 1):
        assert spin in (0.0, 0.99)
        sky = np.ascontiguousarray(sky, dtype=np.uint8)
        cam12 = np.frombuffer(bytes(cam), dtype=np.float32).copy()
        fx_i = np.array([fx.use_bloom, fx.use_vignette, fx.use_ca, fx.use_lens], np.int32)
        fx_f = np.array([fx.bloom_threshold, fx.bloom_intensity, fx.vignette_intensity, fx.ca_amount,
                         fx.distortion_amount], np.float32)
        out = np.zeros((h, w, 4), np.uint8)
        best, mean = C.c_float(), C.c_float()
        rc = self.lib.refcuda_render(1 if spin != 0.0 else 0, w, h, float(time), _ptr(cam12), _ptr(fx_i), _ptr(fx_f),
                                     _ptr(sky), sky.shape[1], sky.shape[0], _ptr(out), int(reps), C.byref(best), C.byref(mean))
        if rc != 0:
            raise RuntimeError(f"refcuda_render rc={rc}")
        return out, float(best.value), float(mean.value)


class RefCudaPlanes:
    """The INSTRUMENTED build of the reference's own CUDA kernel (oracle/ref_cuda_planes_prelude.h force-included in
    front of the unmodified src/raymarcher.cu, same nvcc flags as RefCuda): besides the uchar4 frame it returns the
    kernel's internal locals -- final_hdr, vel, p, intensity, hit_horizon, step count -- at float precision.
    Needs a GPU.  spin must be 0.0 or 0.99 (compile-time there)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_cuda_planes.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        vp = C.c_void_p
        self.lib.refcudap_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp, vp, C.c_int, C.c_int,
                                             vp, vp, vp, vp, vp, vp, vp]
        self.lib.refcudap_render.restype = C.c_int
        self.lib.refcudap_probe.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_float, vp]
        self.lib.refcudap_probe.restype = C.c_int

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libref_cuda_planes.so"))

    def render(self, spin: float, cam, fx, sky: np.ndarray, time: float, w: int, h: int) -> dict:
        """dict(rgba [h,w,4] u8 row-flipped like the reference store; hdr/vel/pos/emis [h,w,4] f32, hit [h,w] u8,
        steps [h,w] i32, all indexed [y][x]); plus derived cls (captured / disk-hit / escaped, SURVEY.md 8a notes:
        disk-hit := not captured and transmittance < 1) and dir = vel / |vel| in float64 -> float32."""
        assert spin in (0.0, 0.99)
        sky = np.ascontiguousarray(sky, dtype=np.uint8)
        cam12 = np.frombuffer(bytes(cam), dtype=np.float32).copy()
        fx_i = np.array([fx.use_bloom, fx.use_vignette, fx.use_ca, fx.use_lens], np.int32)
        fx_f = np.array([fx.bloom_threshold, fx.bloom_intensity, fx.vignette_intensity, fx.ca_amount,
                         fx.distortion_amount], np.float32)
        o = {"rgba": np.zeros((h, w, 4), np.uint8), "hdr": np.zeros((h, w, 4), np.float32),
             "vel": np.zeros((h, w, 4), np.float32), "pos": np.zeros((h, w, 4), np.float32),
             "emis": np.zeros((h, w, 4), np.float32), "hit": np.zeros((h, w), np.uint8), "steps": np.zeros((h, w), np.int32)}
        rc = self.lib.refcudap_render(1 if spin != 0.0 else 0, w, h, float(time), _ptr(cam12), _ptr(fx_i), _ptr(fx_f),
                                      _ptr(sky), sky.shape[1], sky.shape[0], _ptr(o["rgba"]), _ptr(o["hdr"]), _ptr(o["vel"]),
                                      _ptr(o["pos"]), _ptr(o["emis"]), _ptr(o["hit"]), _ptr(o["steps"]))
        if rc != 0:
            raise RuntimeError(f"refcudap_render rc={rc}")
        captured = o["hit"] != 0
        touched = ~captured & (o["hdr"][..., 3] < 1.0)
        o["cls"] = np.where(captured, CLS_CAPTURED, np.where(touched, CLS_DISK_HIT, CLS_ESCAPED)).astype(np.uint8)
        v = o["vel"][..., :3].astype(np.float64)
        n = np.linalg.norm(v, axis=-1, keepdims=True)
        d = np.zeros((h, w, 4), np.float32)
        d[..., :3] = np.where(captured[..., None], 0.0, v / np.maximum(n, 1e-30))
        o["dir"] = d
        return o

    PROBES = {"disk_density": 0, "dust_density": 1, "redshift": 2, "disk_temperature": 3, "noise3d": 4, "fbm5": 5}

    def probe(self, what: str, spin: float, a, b=None, time: float = 0.0) -> np.ndarray:
        """One of the reference's device functions as nvcc compiles it (same flags as the kernel)."""
        assert spin in (0.0, 0.99)
        a = _f32(a)
        n = len(a)
        bb = _f32(b) if b is not None else None
        out = np.empty(n, np.float32)
        rc = self.lib.refcudap_probe(1 if spin != 0.0 else 0, self.PROBES[what], n, _ptr(a), _ptr(bb) if bb is not None else None,
                                     float(time), _ptr(out))
        if rc != 0:
            raise RuntimeError(f"refcudap_probe rc={rc}")
        return out
