"""Float-precision census of OUR planes (both rounding contracts) against the reference's own CUDA arithmetic
(oracle/_ref/libref_cuda_planes.so: src/raymarcher.cu unmodified + instrumentation prelude), on the same GPU.
Also checks that the instrumented build's uchar4 frame equals the un-instrumented libref_cuda.so byte for byte, and
prints the reference-vs-reference figure (reference CUDA build vs reference headers on the host) beside ours.
Usage: python tests/tools/refcuda_planes_census.py [w h] [--host]   (GPU box only; test infrastructure)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relativisticraytracer_b200 as rrt  # noqa: E402
from oracle import Oracle, RefCuda, RefCudaPlanes, available  # noqa: E402
from parity import CAMERAS, census  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
w, h = (int(args[0]), int(args[1])) if len(args) >= 2 else (960, 540)
with_host = "--host" in sys.argv
r = rrt.Renderer(0)
sky_np = rrt.procedural_sky(4096, 2048)
sky = r.create_sky(sky_np)
ref, refp = RefCuda(), RefCudaPlanes()
host = Oracle("reference", auto_build=False) if (with_host and available("reference")) else None
fx = rrt.effects_off()
for cam in ("C0", "C1", "C2", "C3"):
    for spin in (0.0, 0.99):
        cg = rrt.camera_state_from(*CAMERAS[cam])
        plain, _, _ = ref.render(spin, cg, fx, sky_np, 1.0, w, h)
        P = refp.render(spin, cg, fx, sky_np, 1.0, w, h)
        row = {"camera": cam, "spin": spin, "w": w, "h": h,
               "instrumented_bytes_equal_unmodified": bool(np.array_equal(plain, P["rgba"]))}
        for name, flags in (("fmad", 7), ("strict", 3)):
            planes = r.alloc_planes(w, h)
            out = r.render(rrt.default_params(spin_a=spin, flags=flags), cg, fx, sky, 1.0, w, h, planes=planes)
            torch.cuda.synchronize()
            g = {k: v.cpu().numpy() for k, v in planes.items()}
            c = census(P, g)
            touched = (g["cls"] & rrt.CLSF_TOUCHED) != 0
            c["steps_differ"] = int((g["steps"] != P["steps"]).sum())
            c["vel_bits_differ"] = int((g["vel"][..., :3].view(np.uint32) != P["vel"][..., :3].view(np.uint32)).any(axis=-1).sum())
            c["vel_bits_differ_untouched"] = int(((g["vel"][..., :3].view(np.uint32) != P["vel"][..., :3].view(np.uint32)).any(axis=-1) & ~touched).sum())
            c["hdr_bits_differ_untouched"] = int(((g["hdr"][..., :3].view(np.uint32) != P["hdr"][..., :3].view(np.uint32)).any(axis=-1) & ~touched).sum())
            c["touched"] = int(touched.sum())
            rel = np.abs(g["hdr"][..., :3].astype(np.float64) - P["hdr"][..., :3]) / np.maximum(np.abs(P["hdr"][..., :3]), 1e-3)
            c["rgb_over_tol_touched"] = int(((rel.max(axis=-1) > 1e-3) & touched).sum())
            c["rgb_over_tol_untouched"] = int(((rel.max(axis=-1) > 1e-3) & ~touched).sum())
            c["bytes_differ"] = int((np.abs(out.cpu().numpy().astype(int) - plain.astype(int)).max(axis=-1) > 0).sum())
            row[name] = c
        if host is not None:
            f = host.render(host.default_params(spin_a=spin, flags=3), host.camera_from(*CAMERAS[cam]), host.effects_off(), sky_np, 1.0, w, h)
            row["ref_host_vs_ref_cuda"] = census(P, f)
        print(json.dumps(row), flush=True)
# function-level: our FMAD probes vs the reference functions as nvcc compiles them
sys.path.insert(0, os.path.join(ROOT, "tests"))
from inputs import disk_points, phase_space  # noqa: E402
pts = disk_points(seed=31, n=1 << 16)
q, v = phase_space(seed=32, n=1 << 16)
for spin in (0.0, 0.99):
    pf = rrt.default_params(spin_a=spin, flags=7)
    ps = rrt.default_params(spin_a=spin, flags=3)
    for what, ours_f, ours_s, a, b in (
            ("disk_density", r.disk_density(pf, pts, 1.0), r.disk_density(ps, pts, 1.0), pts, None),
            ("dust_density", r.dust_density(pf, pts, 1.0), r.dust_density(ps, pts, 1.0), pts, None),
            ("redshift", r.redshift(pf, q, v), r.redshift(ps, q, v), q, v),
            ("disk_temperature", r.disk_temperature(pf, np.linalg.norm(q, axis=1).astype(np.float32)),
             r.disk_temperature(ps, np.linalg.norm(q, axis=1).astype(np.float32)), np.linalg.norm(q, axis=1).astype(np.float32), None)):
        want = refp.probe(what, spin, a, b, 1.0)
        row = {"probe": what, "spin": spin, "n": int(len(want))}
        for name, got in (("fmad", ours_f), ("strict", ours_s)):
            fin = np.isfinite(want) & np.isfinite(got)
            err = np.abs(got.astype(np.float64) - want)[fin]
            rel = err / np.maximum(np.abs(want[fin]), 1e-6)
            row[name] = {"bits_differ": int((got.view(np.uint32) != want.view(np.uint32)).sum()), "max_abs": float(err.max()),
                         "max_rel": float(rel.max()), "nonfinite_mismatch": int((np.isfinite(want) != np.isfinite(got)).sum())}
        print(json.dumps(row), flush=True)
