"""Skybox sources for the render path.

The reference decodes ``assets/skyboxes/skybox2.jpg`` with stb_image into an RGBA8 equirectangular
map and wraps it in a texture object (src/main.cpp:237-266).  Assets cannot travel with this
repository, so benchmarks and tests use a deterministic procedural map of the same format
(4096x2048 RGBA8 by default): a smooth two-axis gradient (so that bilinear filtering differences stay
far below the parity tolerance) plus a seeded field of small Gaussian "stars".
"""
from __future__ import annotations

import numpy as np


def procedural_sky(width: int = 4096, height: int = 2048, seed: int = 1234, stars: int = 6000) -> np.ndarray:
    """Deterministic RGBA8 equirect sky, shape [height, width, 4]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    v = (np.arange(height, dtype=np.float64) + 0.5) / height
    uu, vv = np.meshgrid(u, v)
    # periodic in u so the wrap seam is smooth
    r = 0.30 + 0.20 * np.sin(2 * np.pi * uu) * np.sin(np.pi * vv)
    g = 0.28 + 0.18 * np.cos(2 * np.pi * uu + 1.0) * np.sin(np.pi * vv) ** 2
    b = 0.40 + 0.25 * np.cos(np.pi * vv) * np.cos(4 * np.pi * uu)
    img = np.stack([r, g, b], axis=-1)
    # stars: 5x5 Gaussian splats at seeded integer positions
    sx = rng.integers(0, width, size=stars)
    sy = rng.integers(2, height - 2, size=stars)
    amp = rng.uniform(0.2, 0.6, size=stars)
    tint = rng.uniform(0.7, 1.0, size=(stars, 3))
    k = np.exp(-0.5 * (np.arange(-2, 3) / 0.9) ** 2)
    k2 = np.outer(k, k)
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            np.add.at(img, (sy + dy, (sx + dx) % width), (amp * k2[dy + 2, dx + 2])[:, None] * tint)
    out = np.empty((height, width, 4), np.uint8)
    out[..., :3] = np.clip(np.floor(img * 255.0 + 0.5), 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def decode_image(path: str) -> np.ndarray:
    """rrt_image_load (C ABI, csrc/rrt_image.cpp): PNG / baseline JPEG -> the RGBA8 array stb_image returns for
    ``stbi_load(path, &w, &h, &c, 4)``, byte for byte.  Raises RrtError for files outside the decoder's subset."""
    import ctypes as C
    from . import _capi
    lib = _capi.load()
    px, w, h = C.POINTER(C.c_uint8)(), C.c_int(), C.c_int()
    rc = lib.rrt_image_load(path.encode(), C.byref(px), C.byref(w), C.byref(h))
    if rc != _capi.OK:
        raise _capi.RrtError(rc, f"{path}: {lib.rrt_image_last_error().decode()}")
    try:
        return np.ctypeslib.as_array(px, shape=(h.value, w.value, 4)).copy()
    finally:
        lib.rrt_image_free(px)


def load_skybox(path: str, width: int | None = None, height: int | None = None) -> np.ndarray:
    """Host half of the reference's ``loadSkybox`` (src/main.cpp:237-245): decode an equirectangular image file
    into the RGBA8, rows-top-down array ``stbi_load(path, &w, &h, &c, 4)`` returns, ready for
    ``Renderer.create_sky`` (the device half, src/main.cpp:246-263).

    * ``.png`` / ``.jpg`` / ``.jpeg``: the library's own decoder (``decode_image`` -> ``rrt_image_load``), whose output
      equals stb_image v2.30's byte for byte -- including for JPEG, where decoders are free to differ (PIL / libjpeg
      do not reproduce stb_image's bytes for the reference's ``skybox2.jpg``).  Files outside its subset (progressive
      JPEG, 16-bit / interlaced PNG) raise instead of being decoded differently from the reference.
    * ``.npy``: an array saved with ``numpy.save`` ([h, w, 3] or [h, w, 4] uint8).
    * ``.rgba`` / ``.raw``: headerless RGBA8, ``width`` and ``height`` required.

    Grey and grey+alpha sources are expanded the way stb_image does for ``req_comp = 4`` (grey replicated to
    RGB, missing alpha = 255)."""
    ext = path.rsplit(".", 1)[-1].lower() if "." in path else ""
    if ext in ("png", "jpg", "jpeg"):
        return decode_image(path)
    if ext == "npy":
        img = np.load(path)
    elif ext in ("rgba", "raw"):
        if not width or not height:
            raise ValueError("raw RGBA skybox needs width and height")
        img = np.fromfile(path, dtype=np.uint8)
        if img.size != width * height * 4:
            raise ValueError(f"{path}: expected {width * height * 4} bytes, found {img.size}")
        img = img.reshape(height, width, 4)
    else:
        raise ValueError(f"{path}: unsupported skybox format (png, jpg, npy, rgba)")
    img = np.asarray(img)
    if img.dtype != np.uint8:
        raise ValueError("skybox must be 8 bits per channel")
    if img.ndim == 2:
        img = img[..., None]
    if img.ndim != 3 or img.shape[2] not in (1, 2, 3, 4):
        raise ValueError(f"unsupported skybox shape {img.shape}")
    h, w, c = img.shape
    out = np.empty((h, w, 4), np.uint8)
    if c >= 3:
        out[..., :3] = img[..., :3]
        out[..., 3] = img[..., 3] if c == 4 else 255
    else:
        out[..., :3] = img[..., :1]
        out[..., 3] = img[..., 1] if c == 2 else 255
    return out
