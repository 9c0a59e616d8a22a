import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import relativisticraytracer_b200 as rrt
from oracle import Oracle
from parity import CAMERAS
r = rrt.Renderer(0); ora = Oracle("port")
sky_np = rrt.procedural_sky(512, 256, seed=1234, stars=0)
sky = r.create_sky(sky_np)
for spin, flags, w, h, fx in ((0.0, 0, 256, 256, "default"), (0.99, 0, 160, 90, "off"), (0.99, 3, 160, 90, "off")):
    pg = rrt.default_params(spin_a=spin, flags=flags); po = ora.default_params(spin_a=spin, flags=flags)
    cg = rrt.camera_state_from(*CAMERAS["C0"])
    fg = rrt.effects_off() if fx == "off" else rrt.default_effects()
    fo = ora.effects_off() if fx == "off" else ora.default_effects()
    planes = r.alloc_planes(w, h)
    r.read_counters(True)
    out = r.render(pg, cg, fg, sky, 1.0, w, h, planes=planes); torch.cuda.synchronize()
    cnt = r.read_counters()
    f = ora.render(po, ora.camera_from(*CAMERAS["C0"]), fo, sky_np, 1.0, w, h)
    g = {k: v.cpu().numpy() for k, v in planes.items()}
    print("config", spin, flags, w, h, "counters gpu", cnt, "ora", f.counters)
    for k in ("cls", "steps", "pos", "vel", "dir"):
        a, b = g[k], getattr(f, k)
        ne = (a != b); ne = ne.any(-1) if ne.ndim == 3 else ne
        print(" ", k, "mismatch pixels:", int(ne.sum()), "first:", np.argwhere(ne)[:5].tolist())
    ne = (g["steps"] != f.steps)
    if ne.any():
        yx = np.argwhere(ne)[0]; y, x = yx
        print("  at", y, x, "steps gpu/ora", g["steps"][y, x], f.steps[y, x], "cls", g["cls"][y, x], f.cls[y, x], "vel", g["vel"][y, x], f.vel[y, x])
