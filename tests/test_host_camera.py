"""Host half of the path (no GPU): the product's camera / keyframe-path code (csrc/rrt_host.cpp behind the C
ABI) against the oracle and the golden vectors generated from the reference."""
import os

import numpy as np

from parity import CAMERAS

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cam_arr(c):
    return np.frombuffer(bytes(c), np.float32)


def test_camera_matches_golden(built):
    import relativisticraytracer_b200 as rrt
    fn = np.load(os.path.join(GOLD, "functions.npz"))
    for i, key in enumerate(("C0", "C1", "C2", "C3")):
        assert np.array_equal(cam_arr(rrt.camera_state_from(*CAMERAS[key])), fn["cameras"][i])


def test_paths_match_golden_and_oracle(built, ora):
    import relativisticraytracer_b200 as rrt
    fn = np.load(os.path.join(GOLD, "functions.npz"))
    assert rrt.path_names() == ["Gargantua Fly-By", "Event Horizon Focus", "Horizon Skimmer"]   # camera_paths.cpp:34,47,60
    assert [rrt.path_duration(i) for i in range(3)] == [25.0, 32.0, 29.0]
    for pi in range(3):
        got = np.stack([cam_arr(rrt.path_state(pi, float(t))[0]) for t in fn["path_t"]])
        assert np.array_equal(got, fn[f"path{pi}"])
    rng = np.random.Generator(np.random.PCG64(21))
    for t in rng.uniform(-3, 40, 500):
        for pi in range(3):
            a, pa = rrt.path_state(pi, t)
            b, pb = ora.path_state(pi, t)
            assert bytes(a) == bytes(b) and np.array_equal(pa, pb)


def test_recording_clock(built):
    """pathTime accumulates 1.0f/24 in float (src/main.cpp:511-516): frame 300 is at 12.5000219, not 12.5"""
    import relativisticraytracer_b200 as rrt
    t = np.float32(0.0)
    for _ in range(300):
        t = np.float32(t + np.float32(1.0) / np.float32(24.0))
    assert rrt.path_clock(300) == float(t)
    assert abs(rrt.path_clock(300) - 12.5000219) < 1e-6
    assert rrt.path_clock(0) == 0.0


def test_bad_path_index(built):
    import pytest
    import relativisticraytracer_b200 as rrt
    with pytest.raises(rrt.RrtError):
        rrt.path_state(3, 1.0)
