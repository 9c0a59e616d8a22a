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
