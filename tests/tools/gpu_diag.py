"""GPU-side diagnostics (run under gpurun): texture-unit calibration dump + parity census per config.
Writes gpurun_out/diag_*.  Uses the oracle as the checker only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import relativisticraytracer_b200 as rrt  # noqa: E402
from oracle import Oracle  # noqa: E402
from parity import CAMERAS, census  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
r = rrt.Renderer(0)
ora = Oracle("port")


def tex_calibration():
    tex = np.zeros((2, 4, 4), np.uint8)
    tex[:, 1::2, :] = 255
    tex[1, :, :] = tex[0, :, :]
    sky = r.create_sky(tex)
    sub = 32
    j = np.arange(-64 * sub, 256 * sub + 64 * sub)
    tx = ((1 + 0.5) / 4 + j / (4.0 * 256 * sub)).astype(np.float32)   # from centre of texel 1 to centre of texel 2
    ty = np.full_like(tx, 0.25)
    hw = r.sky_sample(sky, tx, ty)[:, 0]
    em = ora.tex2d(tex, tx, ty)[:, 0]
    # y direction
    texy = np.zeros((4, 2, 4), np.uint8)
    texy[1::2] = 255
    sky2 = r.create_sky(texy)
    tyy = ((1 + 0.5) / 4 + j / (4.0 * 256 * sub)).astype(np.float32)
    txx = np.full_like(tyy, 0.25)
    hwy = r.sky_sample(sky2, txx, tyy)[:, 0]
    emy = ora.tex2d(texy, txx, tyy)[:, 0]
    # wide texture like the real sky
    big = np.zeros((2048, 4096, 4), np.uint8)
    big[:, 1::2] = 255
    sky3 = r.create_sky(big)
    tx3 = ((1001 + 0.5) / 4096 + j / (4096.0 * 256 * sub)).astype(np.float32)
    hw3 = r.sky_sample(sky3, tx3, np.full_like(tx3, 0.3))[:, 0]
    em3 = ora.tex2d(big, tx3, np.full_like(tx3, 0.3))[:, 0]
    np.savez(os.path.join(OUT, "diag_tex.npz"), j=j, sub=sub, tx=tx, hw=hw, em=em, hwy=hwy, emy=emy, tx3=tx3, hw3=hw3, em3=em3)
    print("tex x: max |hw-em| =", np.abs(hw - em).max(), " y:", np.abs(hwy - emy).max(), " wide:", np.abs(hw3 - em3).max())


def frame_census():
    sky_np = rrt.procedural_sky(512, 256, seed=1234, stars=400)
    smooth = rrt.procedural_sky(512, 256, seed=1234, stars=0)
    res = []
    for name, skyimg in (("stars", sky_np), ("smooth", smooth)):
        sky = r.create_sky(skyimg)
        for cam in ("C0", "C1", "C2", "C3"):
            for spin, flags in ((0.0, 0), (0.99, 0), (0.99, 1), (0.99, 3)):
                w, h = 240, 135
                pg = rrt.default_params(spin_a=spin, flags=flags)
                po = ora.default_params(spin_a=spin, flags=flags)
                cg = rrt.camera_state_from(*CAMERAS[cam])
                planes = r.alloc_planes(w, h)
                r.render(pg, cg, rrt.effects_off(), sky, 1.0, w, h, planes=planes)
                torch.cuda.synchronize()
                g = {k: v.cpu().numpy() for k, v in planes.items()}
                f = ora.render(po, ora.camera_from(*CAMERAS[cam]), ora.effects_off(), skyimg, 1.0, w, h)
                c = census(f, g)
                # emission-only comparison (excludes the texture unit)
                e_rel = np.abs(g["emis"][..., :3].astype(np.float64) - f.emis[..., :3]) / np.maximum(np.abs(f.emis[..., :3]), 1e-3)
                c.update(sky=name, cam=cam, spin=spin, flags=flags, emis_max_rel=float(e_rel.max()),
                         emis_frac_over=float(np.mean(e_rel.max(-1) > 1e-3)),
                         T_max_abs=float(np.abs(g["hdr"][..., 3] - f.hdr[..., 3]).max()),
                         traj_exact=bool(np.array_equal(g["vel"], f.vel) and np.array_equal(g["pos"], f.pos) and np.array_equal(g["steps"], f.steps)))
                res.append(c)
                print(json.dumps(c))
    json.dump(res, open(os.path.join(OUT, "diag_census.json"), "w"), indent=1)


if __name__ == "__main__":
    tex_calibration()
    frame_census()
