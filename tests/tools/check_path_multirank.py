"""Multi-rank check of PathSequence (run under torchrun on N GPUs): the raw stream rank 0 writes equals the frames
rendered one by one on rank 0.  Usage: python -m torch.distributed.run --nproc-per-node N tests/tools/check_path_multirank.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import relativisticraytracer_b200 as rrt  # noqa: E402
from relativisticraytracer_b200.parallel import PathSequence  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, N, PATH = 96, 54, 11, 0
r = rrt.Renderer(local)
sky_np = rrt.procedural_sky(512, 256, seed=1234, stars=400)
sky = r.create_sky(sky_np)
prm, fx = rrt.default_params(spin_a=0.99), rrt.default_effects()
out = "/tmp/path_multirank.rgba"
seq = PathSequence(r, W, H, depth=2)
sink = rrt.FrameSink(out, W, H) if rank == 0 else None
done, launches = seq.render(PATH, N, prm, fx, sky, sink=sink)
if rank == 0:
    sink.close()
    raw = np.fromfile(out, np.uint8).reshape(N, H, W, 4)
    bad = []
    for k in range(1, N + 1):
        t = rrt.path_clock(k, 24.0)
        cam, _ = rrt.path_state(PATH, t)
        want = np.zeros((H, W, 4), np.uint8)
        r.render_host(prm, cam, fx, sky, t, W, H, want)
        if not np.array_equal(raw[k - 1], want):
            bad.append(k)
    print(f"PathSequence world={world}: {done} frames written, mismatching frames: {bad}", flush=True)
    assert done == N and not bad
dist.barrier()
dist.destroy_process_group()
