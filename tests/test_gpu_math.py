"""The kernels evaluate powf / double log10 with two routines of their own (csrc/ig_k_common.cuh): powf_pos, a
transcription of the main path of libdevice's powf (the routine the reference's kernels call, KA:153-163), must be
BIT-IDENTICAL to powf for x > 0; log10_f32 must agree with libdevice's double log10 to < 4e-16 absolute."""
import ctypes as C

import numpy as np
import pytest

from instagraal_b200.synth import WORKLOADS, make_level

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("y", [-1.5, -0.9384134, -1.2345678, -2.9, 2.0, 0.37, -0.0123])
def test_powf_pos_bit_identical_and_log10_accuracy(built, y):
    from instagraal_b200 import _lib as L
    from test_gpu_parity import make_sampler
    s = make_sampler(make_level(WORKLOADS["micro"]))
    out = np.zeros(2, dtype=np.float64)
    for lo, hi in ((1e-4, 1e5), (0.5, 2.0), (1e-30, 1e30)):
        L.check(s._h, L.lib().ig_selftest_math(s._h, 1 << 22, lo, hi, y, out.ctypes.data_as(C.c_void_p)), "ig_selftest_math")
        assert out[0] == 0, ("powf_pos differs from powf", y, lo, hi, out[0])
        assert out[1] < 4e-16 * 40, ("log10_f32", out[1])   # |log10 x| <= 38: absolute error scaled by the exponent term
    s.free_gpu()
