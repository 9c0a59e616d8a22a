"""Renders one frame a few times through the default context (for ncu captures of single launches).
Usage: python tools/render_once.py [--width W --height H --camera C0 --flags 3 --reps 3]  (GPU box; RRT_PIPELINE=fused|split)"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import relativisticraytracer_b200 as rrt  # noqa: E402

CAMS = {"C0": ((0.0, 10.0, -60.0), 0.0, -10.0), "C1": ((15.0, 3.0, -30.0), -26.6, -5.1),
        "C2": ((35.0, 0.8, 10.0), -106.0, -1.2), "C3": ((4.2, 0.6, 4.2), -90.0, -5.7)}
ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--camera", default="C0")
ap.add_argument("--flags", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
r = rrt.Renderer(0)
sky = r.create_sky(rrt.procedural_sky(4096, 2048))
prm = rrt.default_params(spin_a=0.99, flags=a.flags | rrt.FLAG_FMAD)
cam, fx = rrt.camera_state_from(*CAMS[a.camera]), rrt.default_effects()
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r.render(prm, cam, fx, sky, 1.0, a.width, a.height)
    e1.record()
    torch.cuda.synchronize()
    print(f"{a.width}x{a.height} {a.camera} flags={a.flags}: {e0.elapsed_time(e1):.3f} ms", r.split_stats())
