import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (idempotent; make is a no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def ora(built):
    from oracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref(built):
    from oracle import Oracle, available
    if not available("reference"):
        pytest.skip("oracle/_ref/libref_host.so not built (reference tree absent)")
    return Oracle("reference", auto_build=False)


@pytest.fixture(scope="session")
def sky_small():
    from relativisticraytracer_b200 import procedural_sky
    return procedural_sky(512, 256, seed=1234, stars=400)


@pytest.fixture(scope="session")
def sky_smooth():
    """gradient-only sky: neighbouring texels differ by <= 1 count, so the texture unit's 1/256 weight
    quantisation cannot push a pixel past the RGB tolerance whatever libm produced the coordinate"""
    from relativisticraytracer_b200 import procedural_sky
    return procedural_sky(512, 256, seed=1234, stars=0)


@pytest.fixture(scope="session")
def gpu(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import relativisticraytracer_b200 as rrt
    return rrt.Renderer(0)
