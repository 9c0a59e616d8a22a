"""Generate tests/golden/*.npz from the REFERENCE build (oracle/_ref/libref_host.so = the reference's own
headers compiled for the host with -ffp-contract=off).  Run in the container that has /root/reference:

    python tests/tools/make_golden.py

The reference ships no tests or golden vectors; these fixtures are the pin for the oracle port
(tests/test_oracle_golden.py) and travel to the GPU box, where /root/reference does not exist.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from inputs import disk_points, noise_points, phase_space  # noqa: E402
from oracle import Oracle  # noqa: E402
from parity import CAMERAS  # noqa: E402
from relativisticraytracer_b200.skybox import procedural_sky  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
ref = Oracle("reference")

# ---- function-level vectors --------------------------------------------------------------------------
fn = {}
q, v = phase_space(seed=100, n=1024)
fn["ps_q"], fn["ps_v"] = q, v
for spin, tag in ((0.0, "a000"), (0.99, "a099")):
    prm = ref.default_params(spin_a=spin)
    fn[f"acc_{tag}"] = ref.geodesic_acc(prm, q, v)
    for h, ht in ((np.float32(0.3), "h30"), (np.float32(0.3) * np.float32(0.1), "h03"), (np.float32(0.3) * np.float32(0.3), "h09")):
        p1, v1 = ref.rk4_step(prm, q, v, h)
        fn[f"rk4p_{tag}_{ht}"], fn[f"rk4v_{tag}_{ht}"] = p1, v1
    p1, v1 = ref.euler_step(prm, q, v, np.float32(0.3))
    fn[f"eulp_{tag}"], fn[f"eulv_{tag}"] = p1, v1
    fn[f"redshift_{tag}"] = ref.redshift(prm, q, v)
npnt = noise_points(seed=101, n=2048)
fn["noise_p"] = npnt
fn["hash31"] = ref.hash31(npnt)
fn["noise3d"] = ref.noise3d(npnt)
fn["fbm2"] = ref.fbm(npnt, 2)
fn["fbm5"] = ref.fbm(npnt, 5)
dp = disk_points(seed=102, n=2048)
fn["disk_p"] = dp
prm = ref.default_params()
for t, tt in ((0.0, "t0"), (1.0, "t1"), (12.5, "t12")):
    fn[f"disk_density_{tt}"] = ref.disk_density(prm, dp, t)
    fn[f"dust_density_{tt}"] = ref.dust_density(prm, dp, t)
rr = np.linspace(5.0, 40.0, 512).astype(np.float32)
fn["temp_r"], fn["temp"] = rr, ref.disk_temperature(prm, rr)
# host camera + paths
cams = []
for key in ("C0", "C1", "C2", "C3"):
    cams.append(np.frombuffer(bytes(ref.camera_from(*CAMERAS[key])), np.float32))
fn["cameras"] = np.stack(cams)
ts = np.linspace(-1.0, 33.0, 137).astype(np.float32)
fn["path_t"] = ts
for pi in range(3):
    fn[f"path{pi}"] = np.stack([np.frombuffer(bytes(ref.path_state(pi, float(t))[0]), np.float32) for t in ts])
np.savez_compressed(os.path.join(OUT, "functions.npz"), **fn)

# ---- frame planes ----------------------------------------------------------------------------------------
sky = procedural_sky(512, 256, seed=1234, stars=400)
meta = {"sky_sha256": hashlib.sha256(sky.tobytes()).hexdigest(), "sky_args": [512, 256, 1234, 400], "frames": {}}
W, H = 64, 36
fr = {}
for cam in ("C0", "C1", "C2", "C3"):
    for spin, flags, fxname in ((0.0, 0, "off"), (0.99, 3, "off"), (0.99, 1, "default")):
        tag = f"{cam}_a{int(spin * 100):03d}_f{flags}_{fxname}"
        fx = ref.effects_off() if fxname == "off" else ref.default_effects()
        f = ref.render(ref.default_params(spin_a=spin, flags=flags), ref.camera_from(*CAMERAS[cam]), fx, sky, 1.0, W, H)
        for k in ("rgba", "hdr", "dir", "emis", "pos", "vel", "cls", "steps"):
            fr[f"{tag}__{k}"] = getattr(f, k)
        meta["frames"][tag] = {"cam": cam, "spin": spin, "flags": flags, "fx": fxname, "w": W, "h": H, "time": 1.0,
                               "counters": f.counters}
np.savez_compressed(os.path.join(OUT, "frames.npz"), **fr)

# ---- known-answer counters at the survey's sizes -----------------------------------------------------------
ka = {}
f = ref.render(ref.default_params(spin_a=0.0, flags=0), ref.camera_from(*CAMERAS["C0"]), ref.default_effects(), sky, 1.0, 256, 256)
ka["config1_256x256_a0_geodesic_C0_defaultfx"] = f.counters
f = ref.render(ref.default_params(spin_a=0.99, flags=3), ref.camera_from(*CAMERAS["C0"]), ref.default_effects(), sky, 1.0, 480, 270)
ka["480x270_a099_diskdust_C0_defaultfx"] = f.counters
meta["known_answers"] = ka
json.dump(meta, open(os.path.join(OUT, "meta.json"), "w"), indent=1, sort_keys=True)
print("wrote", os.listdir(OUT), {k: os.path.getsize(os.path.join(OUT, k)) for k in os.listdir(OUT)})
print(json.dumps(ka, indent=1))
