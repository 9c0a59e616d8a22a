"""Seeded input generators shared by the parity tests and tests/tools/make_golden.py."""
import numpy as np

RADII = (0.5, 0.99, 1.0, 1.01, 2.0, 2.02, 2.1, 3.0, 6.0, 10.0, 17.99, 18.0, 25.0, 30.0, 60.0, 250.0, 400.0)


def unit_vectors(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v


def phase_space(seed=0, n=4096):
    """positions on shells around the radii the path branches on (+ random radii), random velocities of
    roughly unit length (the loop never renormalises vel, so lengths drift: cover 0.5..1.5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    shells = np.asarray(RADII)[rng.integers(0, len(RADII), size=n)]
    jitter = 1.0 + rng.uniform(-1e-3, 1e-3, size=n)
    rad = np.where(rng.uniform(size=n) < 0.5, shells * jitter, np.exp(rng.uniform(np.log(0.6), np.log(400.0), size=n)))
    q = unit_vectors(rng, n) * rad[:, None]
    v = unit_vectors(rng, n) * rng.uniform(0.5, 1.5, size=(n, 1))
    return q.astype(np.float32), v.astype(np.float32)


def noise_points(seed=1, n=8192):
    """noise lattice coordinates incl. negatives (C fmodf keeps the dividend's sign), exact integers and
    the magnitudes the density functions reach (~ +-700)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.uniform(-700.0, 700.0, size=(n, 3))
    p[: n // 8] = rng.uniform(-4.0, 4.0, size=(n // 8, 3))
    p[n // 8: n // 4] = np.round(rng.uniform(-50.0, 50.0, size=(n // 4 - n // 8, 3)))
    return p.astype(np.float32)


def disk_points(seed=2, n=8192):
    """sample positions inside and around the disk / dust volume (cyl. r in [8,32], |y| < 5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    r = rng.uniform(8.0, 32.0, size=n)
    r[: n // 16] = np.asarray([10.0, 25.0, 21.25, 15.0, 20.0, 9.999, 25.001, 10.001])[rng.integers(0, 8, size=n // 16)]
    phi = rng.uniform(-np.pi, np.pi, size=n)
    y = rng.normal(scale=0.8, size=n)
    y[n // 2:] = rng.normal(scale=0.15, size=n - n // 2)
    return np.stack([r * np.cos(phi), y, r * np.sin(phi)], axis=1).astype(np.float32)
