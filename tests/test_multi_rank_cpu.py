"""The N>1 host logic on CPU: two gloo ranks each fill their packed cyclic band, rank 0 gathers and
assembles, and the result equals the single-rank frame.  (The device half -- rrt_render with a band and
rrt_assemble_bands -- is covered by tests/test_gpu_frames.py::test_bands_equal_full_frame.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pixel(y, x):
    """a stand-in for the pure per-pixel function: depends on (x, y) only"""
    return np.stack([(y * 7 + x) % 251, (y * 3 + x * 5) % 241, (y + x) % 239, np.full_like(y, 255)], axis=-1).astype(np.uint8)


def _worker(rank, world, port, w, h, group, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from relativisticraytracer_b200.parallel import (assemble_host, band_rows_of, frame_owner, gather_bands,
                                                     max_band_rows, path_frames)
    rows = band_rows_of(rank, world, group, h)
    rows_max = max_band_rows(world, group, h)
    packed = np.zeros((rows_max, w, 4), np.uint8)
    yy, xx = np.meshgrid(rows, np.arange(w), indexing="ij")
    packed[: len(rows)] = _pixel(yy, xx)
    got = gather_bands(torch.from_numpy(packed), dst=0)
    ok = True
    if rank == 0:
        frame = assemble_host(got.numpy(), world, group, h)
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        expect = _pixel(yy, xx)[::-1]          # reference store is row-flipped (raymarcher.cu:168)
        ok = bool(np.array_equal(frame, expect))
    else:
        ok = got is None
    # frame-parallel split of a 300-frame path covers every frame exactly once
    mine = path_frames(rank, world, 300)
    t = torch.zeros(301, dtype=torch.int32)
    t[mine] = 1
    dist.all_reduce(t)
    ok = ok and bool((t[1:] == 1).all()) and all(frame_owner(k, world) == rank for k in mine)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_band_gather_and_path_split():
    ctx = mp.get_context("spawn")
    for (w, h, group) in ((40, 37, 8), (16, 64, 1)):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, group, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert sorted(res) == [(0, True), (1, True)]


def test_path_rounds_cover_every_frame_once_in_order():
    """frame-parallel schedule of a camera path (BASELINE config 5): pure host logic"""
    from relativisticraytracer_b200.parallel import frame_owner, path_frames, path_rounds
    for n_frames in (1, 7, 8, 9, 300):
        for nranks in (1, 2, 3, 8):
            for first in (1, 25):
                rounds = path_rounds(n_frames, nranks, first)
                assert all(len(rnd) == nranks for rnd in rounds)
                flat = [f for rnd in rounds for f in sorted(x for x in rnd if x is not None)]
                assert flat == list(range(first, first + n_frames))                 # sink order = frame order
                assert all(f is not None for rnd in rounds[:-1] for f in rnd)     # only the last round may be ragged
                for r in range(nranks):                                            # frame k -> rank k % N
                    col = [rnd[r] for rnd in rounds if rnd[r] is not None]
                    assert all(frame_owner(f, nranks) == r for f in col)
                    if first == 1:
                        assert col == path_frames(r, nranks, n_frames)
