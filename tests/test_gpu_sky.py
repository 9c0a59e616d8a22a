"""Texture-unit calibration: the host emulation in oracle/tex_emul.h (1.8 fixed-point weights, round to
nearest, wrap-x / clamp-y) against the B200 texture unit, through rrt_sky_sample_batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_filter_weight_staircase(gpu, ora):
    tex = np.zeros((2, 4, 4), np.uint8)
    tex[:, 1::2, :] = 255
    sky = gpu.create_sky(tex)
    sub = 32
    j = np.arange(-64 * sub, 320 * sub)
    tx = ((1 + 0.5) / 4 + j / (4.0 * 256 * sub)).astype(np.float32)
    ty = np.full_like(tx, 0.25)
    hw = gpu.sky_sample(sky, tx, ty)
    em = ora.tex2d(tex, tx, ty)
    assert np.abs(hw - em).max() < 2e-5          # 2^-17 residual of the unit's output format
    sky.close()


def test_wrap_and_clamp(gpu, ora, sky_small):
    rng = np.random.Generator(np.random.PCG64(5))
    n = 20000
    tx = rng.uniform(-1.5, 2.5, n).astype(np.float32)     # wrap in x
    ty = rng.uniform(-0.25, 1.25, n).astype(np.float32)   # clamp in y
    tx[:8] = [0.0, 1.0, -1.0, 0.5, 0.99999994, 1.0000001, -1e-8, 2.0]
    ty[:8] = [0.0, 1.0, 0.5, -0.0, 0.99999994, 1.0000001, -1e-8, 0.25]
    sky = gpu.create_sky(sky_small)
    hw = gpu.sky_sample(sky, tx, ty)
    em = ora.tex2d(sky_small, tx, ty)
    # one 1/256 weight step of the largest neighbour contrast bounds any coordinate-rounding disagreement
    err = np.abs(hw - em).max(axis=1)
    # the unit converts coordinates to 1.8 fixed point with its own rounding: a coordinate within float
    # rounding of a weight-bucket edge may land in the neighbouring bucket = one 1/256 step of local contrast
    assert np.quantile(err, 0.9) < 2e-5
    assert err.max() < 1.0 / 256.0
    print("sky emulation: frac within 2e-5 =", float(np.mean(err < 2e-5)), "max =", float(err.max()))
    sky.close()
