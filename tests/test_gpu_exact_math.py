"""The render loop replaces nvcc's guarded IEEE division / square root with their branch-free fast paths
(include/rrt_device.cuh).  This checks, on the device, that they return the correctly rounded result on the
operand domain the loop feeds them: 2^31 random pairs, zero tolerance."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 0xC0FFEE])
def test_fast_div_sqrt_are_correctly_rounded(gpu, seed):
    bad_div, bad_sqrt = gpu.exact_math_selftest(seed=seed, n=1 << 30)
    assert (bad_div, bad_sqrt) == (0, 0)
