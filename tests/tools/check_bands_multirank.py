"""Multi-rank check of the band-parallel frame (run under torchrun on N GPUs): for both exchange paths -- direct peer
stores into rank 0's frame ("peer") and packed bands + NCCL gather + rrt_assemble_bands ("nccl") -- the frame on rank 0
equals the frame one GPU renders alone, byte for byte; same through FramePipeline with several frames in flight and a
host destination.  Usage: python -m torch.distributed.run --nproc-per-node N tests/tools/check_bands_multirank.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relativisticraytracer_b200 as rrt  # noqa: E402
from parity import CAMERAS  # noqa: E402
from relativisticraytracer_b200.parallel import BandedFrame, FramePipeline  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 644, 363          # ragged: 363 rows in groups of 8 over N ranks
r = rrt.Renderer(local)
sky = r.create_sky(rrt.procedural_sky(512, 256, seed=1234, stars=400))
prm, fx = rrt.default_params(spin_a=0.99), rrt.default_effects()
cams = [rrt.camera_state_from(*CAMERAS[c]) for c in ("C0", "C1", "C3")]
want = [r.render(prm, c, fx, sky, 1.0 + i, W, H).clone() for i, c in enumerate(cams)] if rank == 0 else None
torch.cuda.synchronize()
report = []
for exchange in ("peer", "nccl"):
    bf = BandedFrame(r, W, H, 8, exchange=exchange)
    assert bf.exchange == exchange, (bf.exchange, exchange)
    for i, c in enumerate(cams):
        bf.render(prm, c, fx, sky, 1.0 + i)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            if not torch.equal(bf.frame, want[i]):      # say where, before failing
                bad = (bf.frame != want[i]).any(dim=2).any(dim=1).nonzero().flatten().tolist()
                msg = f"BandedFrame {exchange} camera {i}: {len(bad)} rows differ, first {bad[:6]}, frame sum {int(bf.frame.sum())}"
                if bf.peer is not None:
                    chk = torch.zeros_like(want[i])
                    bf.peer.read_into(chk)
                    torch.cuda.synchronize()
                    msg += f"; explicit copy of the peer frame equal to expected: {bool(torch.equal(chk, want[i]))}"
                raise AssertionError(msg)
    for to_host in (False, True):
        pipe = FramePipeline(r, W, H, 8, depth=3, to_host=to_host, exchange=exchange)
        assert pipe.exchange == exchange
        pipe.begin()
        got = []
        for rep in range(3):                       # 9 frames through 3 slots: every slot is reused twice
            for i, c in enumerate(cams):
                pipe.submit(prm, c, fx, sky, 1.0 + i)
                f = pipe.last_frame()              # synchronises on the slot's completion event
                if rank == 0:
                    got.append((i, f.clone() if not to_host else f.cuda()))
        pipe.end()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            for i, f in got:
                assert torch.equal(f, want[i]), f"FramePipeline {exchange} to_host={to_host} camera {i}"
    report.append(exchange)
if rank == 0:
    print(f"bands multirank ok: world={world} exchanges={report}", flush=True)
dist.barrier()
dist.destroy_process_group()
