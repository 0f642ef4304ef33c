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
