"""
CPU: the numba-callable entry points (csr_b200/kernels/cuda_numba.py) compile, and the
nopython -> C-ABI call path works: with no GPU every call must come back as the library's argument
error for an invalid handle, raised from nopython code.
"""

import numpy as np
import pytest

numba = pytest.importorskip("numba")
from numba import njit  # noqa: E402

from csr_b200.kernels import cuda_numba as cn  # noqa: E402


def test_invalid_handle_raises_from_nopython():
    @njit
    def spmv(h, x):
        return cn.mult_vec(h, x)

    with pytest.raises(ValueError, match="invalid cuda kernel handle"):
        spmv(0, np.zeros(3))
    with pytest.raises(ValueError, match="invalid cuda kernel handle"):
        spmv(0, np.zeros(3, np.int32))     # a second specialisation (integer x is promoted to float64)
    with pytest.raises(ValueError):
        cn.mult_ab(0, 0)
    with pytest.raises(ValueError):
        cn.mult_abt(0, 0)
    with pytest.raises(ValueError):
        cn.export_arrays(0)
    with pytest.raises(ValueError):
        cn.dims(0)


def test_release_of_null_handle_is_a_noop_in_nopython():
    @njit
    def f():
        cn.release_handle(0)
        return 7

    assert f() == 7
