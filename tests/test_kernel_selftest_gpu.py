"""Exhaustive GPU self-test of the division-by-constant sequence used by the step kernel."""
import ctypes as C
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_divc_is_ieee_division_for_every_float():
    from taco_b200 import _capi
    L = _capi.lib()
    bad = C.c_uint64(12345)
    _capi.check(L.taco_selftest_divc(0, float(np.float32(0.001)), C.byref(bad)), "taco_selftest_divc")
    assert bad.value == 0, f"{bad.value} (x, C) pairs differ from IEEE x / C: {L.taco_last_error().decode()}"


def test_atan2_poly_within_3_ulp_of_exact():
    """The roll angle's polynomial atan2 (fpv_math.cuh) on 2^22 directions x 9 radii + axes + origin: the accuracy class of the
    libm functions it stands in for (libdevice atan2f: 2 ulp documented; numpy float32 arctan2: 3.3 ulp on random arguments)."""
    from taco_b200 import _capi
    L = _capi.lib()
    worst, bad = C.c_float(-1.0), C.c_uint32(12345)
    _capi.check(L.taco_selftest_atan2(0, 1 << 22, C.byref(worst), C.byref(bad)), "taco_selftest_atan2")
    assert bad.value == 0
    assert 0.0 <= worst.value <= 3.0, worst.value
