"""CPU check of the polynomial atan2 the step kernel uses for the roll angle (fpv_math.cuh:atan2_poly), emulated in numpy float32
(tools/atan2_poly_check.py): the GPU self-test (tests/test_kernel_selftest_gpu.py) runs the device function itself."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from atan2_poly_check import atan2_poly  # noqa: E402


def test_emulated_atan2_poly_accuracy_and_special_values():
    rng = np.random.default_rng(1)
    ang = rng.uniform(-np.pi, np.pi, 400_000); rad = np.exp(rng.uniform(-6, 1, ang.size))
    y = (rad * np.sin(ang)).astype(np.float32); x = (rad * np.cos(ang)).astype(np.float32)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    err = np.abs(atan2_poly(y, x).astype(np.float64) - ref) / ulp
    assert err.max() <= 3.0, err.max()
    f = np.float32
    got = atan2_poly(np.array([0, 1, 0, -1, 0], f), np.array([1, 0, -1, 0, 0], f))
    np.testing.assert_allclose(got, np.array([0, np.pi / 2, np.pi, -np.pi / 2, 0], f), rtol=0, atol=1.2e-7)
