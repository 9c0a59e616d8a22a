"""The render loop replaces nvcc's guarded IEEE division / square root with their branch-free fast paths
(include/rrt_device.cuh).  This checks, on the device, that they return the correctly rounded result on the
operand domain the loop feeds them: 2^31 random pairs, zero tolerance."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 0xC0FFEE])
def test_fast_div_sqrt_are_correctly_rounded(gpu, seed):
    bad_div, bad_sqrt = gpu.exact_math_selftest(seed=seed, n=1 << 30)
    assert (bad_div, bad_sqrt) == (0, 0)


@pytest.mark.parametrize("seed", [3, 0xBADC0DE])
def test_split_powf_equals_libdevice_powf_bit_for_bit(gpu, seed):
    """pow_exp2(pow_log2(x), y) -- libdevice's own powf algorithm cut where the exponent enters, so that one log2 serves the
    three powers of ISCO / r and the two of T / Tref -- against powf itself on 2^29 random positive normal bases, with the
    exponents the media code uses and with random ones: zero tolerance (the noise contrast shaping amplifies any ulp)."""
    bad_used, bad_rand = gpu.exact_pow_selftest(seed=seed, n=1 << 29)
    assert (bad_used, bad_rand) == (0, 0)
