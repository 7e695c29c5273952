"""
CPU: the oracle's restatement of center_rows / unit_rows (csr/transform.py:13-66) against vectors
produced by the unmodified reference (tests/golden/make_golden_normalize.py).

center: bit for bit (float32 and float64).  unit: bit for bit for float32; for float64 the reference's
norm is BLAS dnrm2, the oracle's a plain sqrt(sum of squares), so the tolerance is 4e-15 relative.
"""

import os

import numpy as np
import pytest

from oracle import oracle as orc

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "normalize.npz"))
NAMES = [str(n) for n in Z["names"]]


def mat(name):
    nr, nc, nnz = (int(v) for v in Z[f"{name}.shape"])
    return orc.Mat(nr, nc, nnz, Z[f"{name}.rowptrs"], Z[f"{name}.colinds"], Z[f"{name}.values"])


@pytest.mark.parametrize("name", NAMES)
def test_center_rows_bit_exact(name):
    vec, out = orc.normalize_rows(mat(name), "center")
    assert vec.dtype == Z[f"{name}.values"].dtype
    assert np.array_equal(vec, Z[f"{name}.center.vec"])
    assert np.array_equal(out.values, Z[f"{name}.center.values"])


@pytest.mark.parametrize("name", NAMES)
def test_unit_rows(name):
    with np.errstate(all="ignore"):
        vec, out = orc.normalize_rows(mat(name), "unit")
    gv, gvals = Z[f"{name}.unit.vec"], Z[f"{name}.unit.values"]
    assert vec.dtype == gvals.dtype
    if gvals.dtype == np.float32:
        assert np.array_equal(vec, gv, equal_nan=True) and np.array_equal(out.values, gvals, equal_nan=True)
    else:
        assert np.allclose(vec, gv, rtol=4e-15, atol=0.0, equal_nan=True)
        assert np.allclose(out.values, gvals, rtol=4e-15, atol=0.0, equal_nan=True)


def test_zero_row_becomes_nan_and_empty_rows_stay_zero():
    m = mat("n_f8_zero_row")
    with np.errstate(all="ignore"):
        vec, out = orc.normalize_rows(m, "unit")
    sp, ep = m.rowptrs[3], m.rowptrs[4]
    assert ep > sp and vec[3] == 0.0 and np.all(np.isnan(out.values[sp:ep]))
    e = mat("n_f8_empty_rows")
    vec, _ = orc.normalize_rows(e, "center")
    assert np.all(vec[np.diff(e.rowptrs) == 0] == 0.0)


def test_unknown_normalization_and_missing_values():
    m = mat("n_f8_small")
    with pytest.raises(ValueError):
        orc.normalize_rows(m, "l1")
    with pytest.raises(ValueError):
        orc.normalize_rows(orc.Mat(m.nrows, m.ncols, m.nnz, m.rowptrs, m.colinds, None), "center")
