"""The benchmarked (RRT_FLAG_FMAD) contract pinned to the reference's OWN GPU arithmetic at float precision and at
BASELINE's full sizes (VERDICT r1 "next round" item 1; SURVEY.md 8c golden item 3).

Witness: oracle/_ref/libref_cuda_planes.so = /root/reference/src/raymarcher.cu compiled UNMODIFIED with the flags of
`make ref_cuda` and oracle/ref_cuda_planes_prelude.h force-included, which re-points two call sites by macro so that
the kernel's internal locals -- final_hdr (src/raymarcher.cu:148-150), vel (:129), hit_horizon (:38), transmittance,
intensity_*, the number of integrate_rk4 calls (:64) -- are also written to planes.  Nothing of the kernel is restated.

Tolerances are north_star's (class exact, direction 1e-5 rad, linear RGB 1e-3 relative); what is measured on a B200
(tests/tools/refcuda_planes_census.py, profiles/r2_refcuda_planes_census.md) is far tighter and is asserted here too:
every exit velocity and step count bit-identical on every pixel of every camera, final_hdr bit-identical on every pixel
whose ray never touched a medium and within 1e-6 relative on the others.  The strict contract sits, as designed, at the
reference-vs-reference distance (host headers vs CUDA build) from this witness; that is asserted as well.
"""
import numpy as np
import pytest

from inputs import disk_points, phase_space
from parity import CAMERAS, DIR_TOL_RAD, RGB_TOL_REL, census

pytestmark = pytest.mark.gpu

FMAD = 4


@pytest.fixture(scope="module")
def refp():
    from oracle import RefCudaPlanes
    if not RefCudaPlanes.available():
        pytest.skip("oracle/_ref/libref_cuda_planes.so not built (reference tree absent at build time)")
    return RefCudaPlanes()


@pytest.fixture(scope="module")
def refcuda():
    from oracle import RefCuda
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    return RefCuda()


@pytest.fixture(scope="module")
def sky_big():
    import relativisticraytracer_b200 as rrt
    return rrt.procedural_sky(4096, 2048)


def _ours(gpu, sky_np, cam, spin, flags, w, h, fx):
    import relativisticraytracer_b200 as rrt
    import torch
    sky = gpu.create_sky(sky_np)
    planes = gpu.alloc_planes(w, h)
    out = gpu.render(rrt.default_params(spin_a=spin, flags=flags), rrt.camera_state_from(*CAMERAS[cam]), fx, sky, 1.0, w, h, planes=planes)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in planes.items()}
    g["rgba"] = out.cpu().numpy()
    sky.close()
    return g


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _check_fmad_against_reference_cuda(g, P, plain_rgba, tag):
    import relativisticraytracer_b200 as rrt
    touched = (g["cls"] & rrt.CLSF_TOUCHED) != 0
    # termination class, step count: exact on every pixel
    assert np.array_equal(g["cls"] & rrt.CLS_MASK, P["cls"]), f"{tag}: termination class differs from the reference CUDA build"
    assert np.array_equal(g["steps"], P["steps"]), f"{tag}: step counts differ"
    # exit state: bit-identical on every pixel (the whole trajectory is the reference's, operation for operation)
    assert np.array_equal(_bits(g["vel"][..., :3]), _bits(P["vel"][..., :3])), f"{tag}: exit velocity not bit-identical"
    assert np.array_equal(_bits(g["pos"][..., :3]), _bits(P["pos"][..., :3])), f"{tag}: exit position not bit-identical"
    c = census(P, g)
    assert c["class_flips"] == 0 and c["dir_max_rad"] <= 1e-6 < DIR_TOL_RAD, (tag, c)
    # linear RGB (final_hdr, effects off): bit-identical where no medium was touched, 1e-6 where one was
    same = (_bits(g["hdr"][..., :3]) == _bits(P["hdr"][..., :3])).all(axis=-1)
    assert same[~touched].all(), f"{tag}: {int((~same[~touched]).sum())} untouched pixels differ in final_hdr"
    assert c["rgb_max_rel"] <= 1e-6 < RGB_TOL_REL, (tag, c)
    assert np.array_equal(_bits(g["hdr"][..., 3]), _bits(P["hdr"][..., 3])) or \
        np.abs(g["hdr"][..., 3] - P["hdr"][..., 3]).max() <= 1e-6, f"{tag}: transmittance"
    # 8-bit frame against the UNMODIFIED kernel: identical except a handful of touched pixels one count off
    d = np.abs(g["rgba"].astype(int) - plain_rgba.astype(int)).max(axis=-1)[::-1]     # frame rows are flipped (:168)
    assert d[~touched].max() == 0, f"{tag}: untouched pixels differ from the unmodified reference kernel"
    assert d.max() <= 1 and (d > 0).sum() <= max(3, g["cls"].size // 200000), f"{tag}: {(d > 0).sum()} bytes differ, max {d.max()}"
    return c


@pytest.mark.parametrize("cam", ["C0", "C1", "C2", "C3"])
@pytest.mark.parametrize("spin", [0.0, 0.99])
def test_instrumentation_does_not_change_the_reference_kernel(gpu, refp, refcuda, sky_small, cam, spin):
    """The instrumented build's uchar4 frame against the un-instrumented libref_cuda.so: byte-identical on every pixel
    whose ray touched no medium; observing intensity_* / transmittance lets nvcc fuse the emission accumulation of a
    touched pixel differently in the last bit now and then (measured: <= 7 of 2 M pixels, one 8-bit count)."""
    import relativisticraytracer_b200 as rrt
    w, h = 480, 270
    cg, fx = rrt.camera_state_from(*CAMERAS[cam]), rrt.default_effects()
    plain, _, _ = refcuda.render(spin, cg, fx, sky_small, 1.0, w, h)
    P = refp.render(spin, cg, fx, sky_small, 1.0, w, h)
    d = np.abs(plain.astype(int) - P["rgba"].astype(int)).max(axis=-1)[::-1]
    touched = (P["hdr"][..., 3] < 1.0) | (P["emis"][..., :3] != 0).any(axis=-1)
    assert d[~touched].max() == 0
    assert d.max() <= 1 and (d > 0).sum() <= 3
    assert int(P["steps"].max()) <= 2000 and int(P["steps"].min()) >= 0


@pytest.mark.parametrize("cam", ["C0", "C1", "C2", "C3"])
@pytest.mark.parametrize("spin", [0.0, 0.99])
def test_fmad_planes_equal_reference_cuda_arithmetic(gpu, refp, refcuda, sky_small, cam, spin):
    import relativisticraytracer_b200 as rrt
    w, h = 480, 270
    cg, fx = rrt.camera_state_from(*CAMERAS[cam]), rrt.effects_off()
    plain, _, _ = refcuda.render(spin, cg, fx, sky_small, 1.0, w, h)
    P = refp.render(spin, cg, fx, sky_small, 1.0, w, h)
    g = _ours(gpu, sky_small, cam, spin, 3 | FMAD, w, h, fx)
    _check_fmad_against_reference_cuda(g, P, plain, f"{cam} a={spin} {w}x{h}")


@pytest.mark.parametrize("w,h,cam", [(1920, 1080, "C0"), (1920, 1080, "C1"), (3840, 2160, "C0")])
def test_fmad_full_size_frames_equal_reference_cuda(gpu, refp, refcuda, sky_big, w, h, cam):
    """BASELINE configs 3 (1080p disk + dust) and 4 (the 4K bench frame), a = 0.99, against the reference's own kernel on
    the same GPU: planes at float precision and the 8-bit frame with the reference's default effects."""
    import relativisticraytracer_b200 as rrt
    cg = rrt.camera_state_from(*CAMERAS[cam])
    fx = rrt.effects_off()
    plain, _, _ = refcuda.render(0.99, cg, fx, sky_big, 1.0, w, h)
    P = refp.render(0.99, cg, fx, sky_big, 1.0, w, h)
    g = _ours(gpu, sky_big, cam, 0.99, 3 | FMAD, w, h, fx)
    _check_fmad_against_reference_cuda(g, P, plain, f"{cam} {w}x{h}")
    # and the frame as the viewer shows it (bloom, vignette, lens distortion on: reference defaults)
    fxd = rrt.default_effects()
    plain_d, _, _ = refcuda.render(0.99, cg, fxd, sky_big, 1.0, w, h)
    gd = _ours(gpu, sky_big, cam, 0.99, 3 | FMAD, w, h, fxd)
    touched = (gd["cls"] & rrt.CLSF_TOUCHED) != 0
    d = np.abs(gd["rgba"].astype(int) - plain_d.astype(int)).max(axis=-1)[::-1]
    assert d[~touched].max() == 0
    assert d.max() <= 1 and (d > 0).sum() <= max(3, w * h // 200000)


def test_strict_1080p_frame_against_reference_headers_on_host(gpu, ref, sky_smooth):
    """The strict contract at BASELINE's 1080p size against oracle/_ref/libref_host.so (the reference's unmodified
    headers, -ffp-contract=off): trajectory bit-identical on every pixel, linear RGB within north_star's 1e-3."""
    import relativisticraytracer_b200 as rrt
    w, h = 1920, 1080
    fx = rrt.effects_off()
    g = _ours(gpu, sky_smooth, "C0", 0.99, 3, w, h, fx)
    f = ref.render(ref.default_params(spin_a=0.99, flags=3), ref.camera_from(*CAMERAS["C0"]), ref.effects_off(), sky_smooth, 1.0, w, h)
    for k in ("cls", "steps", "pos", "vel", "dir"):
        assert np.array_equal(g[k], getattr(f, k)), k
    c = census(f, g)
    assert c["class_flips"] == 0 and c["dir_max_rad"] == 0.0
    assert c["rgb_frac_over_tol"] == 0.0, c


def test_strict_contract_sits_at_the_reference_vs_reference_distance(gpu, refp, ref, sky_small):
    """Context for the two contracts: against the reference's CUDA arithmetic the strict contract differs exactly as much
    as the reference's own headers compiled for a host do (same census), i.e. no implementation can be bit-identical
    to both roundings of the reference; the FMAD contract is the one that matches its GPU build."""
    import relativisticraytracer_b200 as rrt
    w, h = 480, 270
    cg, fx = rrt.camera_state_from(*CAMERAS["C1"]), rrt.effects_off()
    P = refp.render(0.99, cg, fx, sky_small, 1.0, w, h)
    g = _ours(gpu, sky_small, "C1", 0.99, 3, w, h, fx)
    f = ref.render(ref.default_params(spin_a=0.99, flags=3), ref.camera_from(*CAMERAS["C1"]), ref.effects_off(), sky_small, 1.0, w, h)
    ours, refs = census(P, g), census(P, f)
    assert ours["class_flips"] == refs["class_flips"]
    assert abs(ours["dir_frac_over_tol"] - refs["dir_frac_over_tol"]) < 1e-6
    assert abs(ours["rgb_frac_over_tol"] - refs["rgb_frac_over_tol"]) < 2e-3
    assert refs["rgb_frac_over_tol"] > 1e-3      # the two roundings of the reference itself do not meet 1e-3 everywhere


# ---- function-level: the media functions in the FMAD unit against the reference functions as nvcc compiles them ----
def _away_from_gates(pts, eps=2e-3):
    """The density functions switch on cylindrical radius gates (ISCO 10, DISK_OUT 25, taper 21.25, smoothstep edges 15, 20);
    a sample whose radius rounds to the other side of a gate in one build is a different branch, not a rounding error."""
    r = np.sqrt(pts[:, 0].astype(np.float64) ** 2 + pts[:, 2].astype(np.float64) ** 2)
    ok = np.ones(len(pts), bool)
    for gate in (10.0, 25.0, 21.25, 15.0, 20.0):
        ok &= np.abs(r - gate) > eps
    return ok


@pytest.mark.parametrize("spin", [0.0, 0.99])
def test_fmad_media_probes_against_nvcc_compiled_reference_functions(gpu, refp, spin):
    """getAccretionDensity / getDustCloudDensity / calculateRedshiftFactor / getDiskTemperature of the FMAD unit vs the
    reference's functions compiled by nvcc with its default -fmad=true (a probe kernel in libref_cuda_planes.so).
    nvcc fuses a function differently standalone than inlined into raymarch_kernel, so this pair is close, not
    bit-identical (the bit-level pin of the media path is the whole-frame test above): the noise hash extracts the
    fraction of numbers ~1e4, so one ulp in its argument moves a lattice value by ~1e-3 and the contrast shaping
    (pow 1.6 / pow 4, smoothstep) amplifies that on a small share of samples."""
    import relativisticraytracer_b200 as rrt
    prm = rrt.default_params(spin_a=spin, flags=3 | FMAD)
    pts = disk_points(seed=31, n=1 << 15)
    ok = _away_from_gates(pts)
    for what, ours in (("disk_density", gpu.disk_density(prm, pts, 1.0)), ("dust_density", gpu.dust_density(prm, pts, 1.0))):
        want = refp.probe(what, spin, pts, None, 1.0)
        assert np.array_equal(ours[ok] == 0.0, want[ok] == 0.0), f"{what}: zero / non-zero pattern differs away from the gates"
        err = np.abs(ours.astype(np.float64) - want)[ok]
        scale = np.maximum(np.abs(want[ok]), 1e-2)
        rel = err / scale
        assert np.quantile(rel, 0.99) < 1e-4, (what, float(np.quantile(rel, 0.99)))
        assert np.quantile(rel, 0.999) < 5e-3, (what, float(np.quantile(rel, 0.999)))
        assert np.mean(ours[ok].view(np.uint32) == want[ok].view(np.uint32)) > 0.9, what
    q, v = phase_space(seed=32, n=1 << 15)
    g_ours, g_want = gpu.redshift(prm, q, v), refp.probe("redshift", spin, q, v, 0.0)
    fin = np.isfinite(g_want)
    assert np.array_equal(fin, np.isfinite(g_ours))
    assert np.abs(g_ours[fin] - g_want[fin]).max() <= 2e-5 * np.maximum(np.abs(g_want[fin]), 1.0).max()
    np.testing.assert_allclose(g_ours[fin], g_want[fin], rtol=2e-5, atol=1e-6)
    r = np.linalg.norm(q, axis=1).astype(np.float32)
    assert np.array_equal(gpu.disk_temperature(prm, r).view(np.uint32), refp.probe("disk_temperature", spin, r).view(np.uint32))
