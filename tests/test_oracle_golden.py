"""The oracle port (oracle/rrt_oracle.c) against the committed golden vectors, which tests/tools/make_golden.py
generated from the reference's own headers (oracle/_ref).  Bit-for-bit: same libm, same operations."""
import hashlib
import json
import os

import numpy as np
import pytest

from parity import CAMERAS, census

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def fn():
    return np.load(os.path.join(GOLD, "functions.npz"))


@pytest.fixture(scope="module")
def frames():
    return np.load(os.path.join(GOLD, "frames.npz")), json.load(open(os.path.join(GOLD, "meta.json")))


def same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint8), np.asarray(b).view(np.uint8))


@pytest.mark.parametrize("spin,tag", [(0.0, "a000"), (0.99, "a099")])
def test_geodesic_functions(ora, fn, spin, tag):
    prm = ora.default_params(spin_a=spin)
    q, v = fn["ps_q"], fn["ps_v"]
    assert same(ora.geodesic_acc(prm, q, v), fn[f"acc_{tag}"])
    for h, ht in ((np.float32(0.3), "h30"), (np.float32(0.3) * np.float32(0.1), "h03"), (np.float32(0.3) * np.float32(0.3), "h09")):
        p1, v1 = ora.rk4_step(prm, q, v, h)
        assert same(p1, fn[f"rk4p_{tag}_{ht}"]) and same(v1, fn[f"rk4v_{tag}_{ht}"])
    p1, v1 = ora.euler_step(prm, q, v, np.float32(0.3))
    assert same(p1, fn[f"eulp_{tag}"]) and same(v1, fn[f"eulv_{tag}"])
    assert same(ora.redshift(prm, q, v), fn[f"redshift_{tag}"])


def test_noise_functions(ora, fn):
    p = fn["noise_p"]
    assert same(ora.hash31(p), fn["hash31"])
    assert same(ora.noise3d(p), fn["noise3d"])
    assert same(ora.fbm(p, 2), fn["fbm2"])
    assert same(ora.fbm(p, 5), fn["fbm5"])


def test_density_functions(ora, fn):
    prm = ora.default_params()
    for t, tt in ((0.0, "t0"), (1.0, "t1"), (12.5, "t12")):
        assert same(ora.disk_density(prm, fn["disk_p"], t), fn[f"disk_density_{tt}"])
        assert same(ora.dust_density(prm, fn["disk_p"], t), fn[f"dust_density_{tt}"])
    assert same(ora.disk_temperature(prm, fn["temp_r"]), fn["temp"])


def test_host_camera_and_paths(ora, fn):
    for i, key in enumerate(("C0", "C1", "C2", "C3")):
        assert same(np.frombuffer(bytes(ora.camera_from(*CAMERAS[key])), np.float32), fn["cameras"][i])
    for pi in range(3):
        got = np.stack([np.frombuffer(bytes(ora.path_state(pi, float(t))[0]), np.float32) for t in fn["path_t"]])
        assert same(got, fn[f"path{pi}"])


def test_frames(ora, frames, sky_small):
    npz, meta = frames
    assert hashlib.sha256(sky_small.tobytes()).hexdigest() == meta["sky_sha256"], "procedural sky changed"
    for tag, m in meta["frames"].items():
        fx = ora.effects_off() if m["fx"] == "off" else ora.default_effects()
        f = ora.render(ora.default_params(spin_a=m["spin"], flags=m["flags"]), ora.camera_from(*CAMERAS[m["cam"]]), fx,
                       sky_small, m["time"], m["w"], m["h"])
        for k in ("rgba", "hdr", "dir", "emis", "pos", "vel", "cls", "steps"):
            assert same(getattr(f, k), npz[f"{tag}__{k}"]), (tag, k)
        assert f.counters == m["counters"], tag


def test_known_answers_config1(ora, frames, sky_small):
    """BASELINE config 1 at full size on the CPU: 256x256, a=0, geodesic only."""
    _, meta = frames
    f = ora.render(ora.default_params(spin_a=0.0, flags=0), ora.camera_from(*CAMERAS["C0"]), ora.default_effects(),
                   sky_small, 1.0, 256, 256)
    assert f.counters == meta["known_answers"]["config1_256x256_a0_geodesic_C0_defaultfx"]
    assert f.counters["rk4_steps"] == 69585851
    assert (f.counters["n_captured"], f.counters["n_escaped"], f.counters["n_exhausted"]) == (374, 63208, 1954)


def test_row_ranges_and_empty(ora, sky_small):
    """rows [y0,y1) only touch their own pixels; an empty range is a no-op; bad ranges are rejected."""
    prm, cam, fx = ora.default_params(spin_a=0.99), ora.camera_from(*CAMERAS["C1"]), ora.effects_off()
    full = ora.render(prm, cam, fx, sky_small, 1.0, 48, 27)
    part = ora.render(prm, cam, fx, sky_small, 1.0, 48, 27, y0=5, y1=11)
    assert same(part.hdr[5:11], full.hdr[5:11]) and not part.hdr[:5].any() and not part.hdr[11:].any()
    empty = ora.render(prm, cam, fx, sky_small, 1.0, 48, 27, y0=7, y1=7)
    assert empty.counters["rk4_steps"] == 0
    with pytest.raises(ValueError):
        ora.render(prm, cam, fx, sky_small, 1.0, 48, 27, y0=0, y1=28)


@pytest.mark.parametrize("cam", ["C0", "C1", "C2", "C3"])
@pytest.mark.parametrize("tag,spin", [("a000", 0.0), ("a099", 0.99)])
def test_fmad_twin_against_reference_cuda_frames(ora, sky_small, cam, tag, spin):
    """The CPU twin of the FMAD contract (oracle port with ORA_FLAG_FMAD) against uchar4 frames rendered on a B200
    by the reference's OWN CUDA kernel (tests/golden/refcuda_frames.npz, tests/tools/make_golden_refcuda.py): every byte
    within one count (what host libm vs libdevice atan2f/asinf/expf and the emulated texture filter leave), on
    well under 1 % of the pixels -- whereas the unfused arithmetic is off by up to tens of counts where rays touch
    the media, because the noise hash amplifies the rounding difference of an unfused dot product."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "refcuda_frames.npz")
    ref = np.load(path)[f"{cam}_{tag}"]
    h, w = ref.shape[:2]
    fused = ora.render(ora.default_params(spin_a=spin, flags=3 | 4), ora.camera_from(*CAMERAS[cam]), ora.default_effects(),
                       sky_small, 1.0, w, h)
    d = np.abs(fused.rgba.astype(int) - ref.astype(int)).max(axis=-1)
    assert d.max() <= 1
    assert (d > 0).mean() < 0.01
    if cam != "C0":   # enough medium in view for the hash sensitivity to show
        unfused = ora.render(ora.default_params(spin_a=spin, flags=3), ora.camera_from(*CAMERAS[cam]), ora.default_effects(),
                             sky_small, 1.0, w, h)
        du = np.abs(unfused.rgba.astype(int) - ref.astype(int)).max(axis=-1)
        assert du.max() >= 2 and (du > 0).sum() > (d > 0).sum()


@pytest.mark.parametrize("cam", ["C0", "C3"])
def test_the_two_rounding_contracts_agree_within_north_star_tolerances(ora, sky_smooth, cam):
    """The reference's numbers exist in two roundings (its CUDA build fuses a*b+c, its headers on a host need not):
    their mutual distance is the floor under any parity claim.  On the CPU twins: no termination
    class flips, exit directions within 1e-5 rad except on a fraction of a percent of strongly lensed rays, linear
    RGB within 1e-3 except where the noise hash amplifies the rounding of an unfused dot product."""
    w, h = 160, 90
    args = (ora.camera_from(*CAMERAS[cam]), ora.effects_off(), sky_smooth, 1.0, w, h)
    a = ora.render(ora.default_params(spin_a=0.99, flags=3), *args)
    b = ora.render(ora.default_params(spin_a=0.99, flags=3 | 4), *args)
    c = census(a, b)
    assert c["class_flips"] == 0
    assert c["dir_frac_over_tol"] < 0.02 and c["dir_max_rad"] < 1e-2
    assert c["rgb_frac_over_tol"] < 0.05
    assert not np.array_equal(a.vel, b.vel)            # they are different roundings
    assert abs(int(a.counters["rk4_steps"]) - int(b.counters["rk4_steps"])) < 2000
