"""World-size > 1 on real GPUs, under the same launcher the driver uses (torch.distributed.run, one rank per GPU, NCCL).
Skipped on a box with a single GPU; on the multi-GPU box they cover BASELINE configs 4 and 5 end to end:
row-band frames (trace -> NCCL gather / peer stores -> assembled frame on rank 0) and the frame-parallel camera path."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(script, nproc, port, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script), *args]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, f"{' '.join(cmd)}\n{res.stdout[-3000:]}\n{res.stderr[-3000:]}"
    return res.stdout


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_path_sequence_frame_parallel_matches_single_gpu_frames(gpu, nproc):
    """BASELINE config 5 (frame k on rank k mod N, gathered to rank 0 in frame order): the stream rank 0 writes equals
    the frames rendered one at a time (tests/tools/check_path_multirank.py asserts it on rank 0)."""
    if _ngpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = _torchrun("tests/tools/check_path_multirank.py", nproc, 29541 + nproc)
    assert "mismatching frames: []" in out


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_banded_frame_multirank_equals_single_gpu(gpu, nproc):
    """BASELINE config 4: a frame cut into cyclic row bands over N GPUs is byte-identical to the single-GPU frame, through
    both exchange paths (NCCL gather + rrt_assemble_bands, and direct peer stores into rank 0's frame)."""
    if _ngpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = _torchrun("tests/tools/check_bands_multirank.py", nproc, 29561 + nproc)
    assert "bands multirank ok" in out
