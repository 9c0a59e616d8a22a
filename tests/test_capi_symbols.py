"""The C-ABI library loads on a CPU-only box and exports every symbol include/rrt.h declares, plus the
reference-mangled launch_raymarch.  No compute entry point is called here (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rrt.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rrt_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    from relativisticraytracer_b200 import _capi
    lib = _capi.load()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rrt.h but not exported"
    assert sorted(_capi.SYMBOLS) == names, "python binding list and header disagree"


def test_reference_launcher_symbol(built):
    """same mangled name as the reference's launch_raymarch (include/raymarcher.h:19), SURVEY.md 2 row 2"""
    from relativisticraytracer_b200 import _capi
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "_Z15launch_raymarchP6uchar4iif11CameraStatey13CameraEffects" in out
    assert "rrt_compat_set_params" in out


def test_struct_layouts_and_defaults(built):
    from relativisticraytracer_b200 import _capi
    lib = _capi.load()
    assert lib.rrt_abi_version() == 1
    assert b"sm_100a" in lib.rrt_build_info() and b"fmad=false" in lib.rrt_build_info()
    assert C.sizeof(_capi.Camera) == 48 and C.sizeof(_capi.Effects) == 36 and C.sizeof(_capi.Params) == 64
    assert C.sizeof(_capi.Counters) == 64 and C.sizeof(_capi.Planes) == 56 and C.sizeof(_capi.Band) == 12
    p = _capi.default_params()
    # include/config.h
    assert (p.spin_a, p.event_horizon, p.isco_radius, p.disk_out, p.max_steps) == (0.0, 2.0, 10.0, 25.0, 2000)
    assert abs(p.step_size - 0.3) < 1e-7 and abs(p.disk_temp_ref - 1.5e7) < 1 and p.flags == 7   # disk | dust | RRT_FLAG_FMAD
    e = _capi.default_effects()
    # camera_settings.h:5-16
    assert (e.use_bloom, e.use_vignette, e.use_ca, e.use_lens) == (1, 1, 0, 1)
    assert abs(e.bloom_threshold - 0.8) < 1e-7 and abs(e.distortion_amount - 0.15) < 1e-7


def test_no_device_is_a_loud_error(built):
    """without a GPU the library refuses to create a context instead of falling back to anything"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from relativisticraytracer_b200 import _capi
    lib = _capi.load()
    ctx = C.c_void_p()
    rc = lib.rrt_context_create(0, C.byref(ctx))
    assert rc == _capi.ERR_NO_DEVICE and not ctx.value
    assert b"no CUDA device" in lib.rrt_last_error(None)
    import relativisticraytracer_b200 as rrt
    with pytest.raises(RuntimeError):
        rrt.Renderer()


def test_band_rows_host_logic(built):
    from relativisticraytracer_b200 import _capi
    from relativisticraytracer_b200.parallel import band_rows_of, max_band_rows
    lib = _capi.load()
    for h in (1, 7, 117, 1080, 2160):
        for nranks in (1, 2, 3, 4, 8):
            for group in (1, 5, 8, 16):
                tot = 0
                seen = set()
                for r in range(nranks):
                    b = _capi.Band(r, nranks, group)
                    n = lib.rrt_band_rows(C.byref(b), h)
                    rows = band_rows_of(r, nranks, group, h)
                    assert n == len(rows)
                    seen.update(rows.tolist())
                    tot += n
                assert tot == h and seen == set(range(h))
                assert max_band_rows(nranks, group, h) >= (h + nranks - 1) // nranks
    assert lib.rrt_band_rows(C.byref(_capi.Band(2, 2, 8)), 100) == _capi.ERR_BAD_ARG
    assert lib.rrt_band_rows(None, 100) == 100


def test_header_is_plain_c(built, tmp_path):
    """include/rrt.h is the FFI surface: it must compile as strict C11 (no C++-isms, no CUDA types) and link against
    the library from a C translation unit."""
    src = tmp_path / "use_rrt.c"
    src.write_text(
        '#include "rrt.h"\n'
        "#include <stdio.h>\n"
        "int main(void) {\n"
        "    rrt_params p; rrt_effects e; rrt_camera c; float pos[3] = {0.0f, 10.0f, -60.0f};\n"
        "    rrt_default_params(&p); rrt_default_effects(&e); rrt_camera_from(pos, 0.0f, -10.0f, &c);\n"
        '    printf("%d %d %u %.3f %d %s\\n", rrt_abi_version(), p.max_steps, p.flags, c.forward[2], rrt_path_count(), rrt_path_name(0));\n'
        "    return sizeof(rrt_camera) == 48 && sizeof(rrt_effects) == 36 && sizeof(rrt_params) == 64 ? 0 : 1;\n"
        "}\n")
    from relativisticraytracer_b200 import _capi
    pkg = os.path.dirname(_capi.LIB_PATH)
    exe = str(tmp_path / "use_rrt")
    res = subprocess.run(["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                          str(src), "-o", exe, "-L" + pkg, "-lrrt_b200", "-Wl,-rpath," + pkg], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0
    assert out.stdout.split()[:3] == ["1", "2000", "7"] and "Gargantua" in out.stdout
