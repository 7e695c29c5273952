"""
GPU: property tests, re-statements of the reference's own hot-path tests with the
``cuda`` kernel in the ``kernel`` fixture (reference tests/test_multiply.py,
test_mult_vec.py, test_handles.py, test_transform.py:77-87, test_transpose.py),
plus structure parity against the oracle on every drawn input.
"""

import numpy as np
import pytest
import hypothesis.strategies as st
import hypothesis.extra.numpy as nph
from hypothesis import given, assume
from pytest import approx

from csr_b200 import CSR
from oracle import oracle as orc
from util import canonical, assert_values_close, value_tol, spgemm_terms, spmv_terms

pytestmark = pytest.mark.gpu


@st.composite
def finite_arrays(draw, shape, dtype=np.float64, min_value=-1.0e3, max_value=1.0e3, **kwargs):
    "csr/test_utils.py:22-27"
    dtype = np.dtype(dtype)
    elts = nph.from_dtype(dtype, min_value=min_value, max_value=max_value,
                          allow_infinity=False, allow_nan=False, **kwargs)
    return draw(nph.arrays(dtype, shape, elements=elts))


@st.composite
def csrs(draw, nrows=None, ncols=None, max_density=0.5, values=None, dtype=('f4', 'f8')):
    "csr/test_utils.py:30-74: COO draws in random order, explicit zeros removed."
    ncols = draw(st.integers(1, 80)) if ncols is None else ncols
    nrows = draw(st.integers(1, 80)) if nrows is None else nrows
    nnz_ub = int(np.ceil(nrows * ncols * max_density))
    nnz = draw(st.integers(0, nnz_ub))
    coords = draw(nph.arrays(np.int32, nnz, elements=st.integers(0, nrows * ncols - 1), unique=True))
    rows = np.mod(coords, nrows, dtype=np.int32)
    cols = np.floor_divide(coords, nrows, dtype=np.int32)
    dt = np.dtype(draw(st.sampled_from(list(dtype))) if not isinstance(dtype, str) else dtype)
    if values is None:
        values = draw(st.booleans())
    if values:
        vals = draw(finite_arrays(nnz, dtype=dt))
        nz = vals != 0.0
        rows, cols, vals = rows[nz], cols[nz], vals[nz]
    else:
        vals = None
    return CSR.from_coo(rows, cols, vals, (nrows, ncols))


@st.composite
def mm_pairs(draw, max_shape=(100, 100, 100), **kw):
    "csr/test_utils.py:86-101"
    mr, mm, mc = max_shape
    rows, mids, cols = draw(st.integers(1, mr)), draw(st.integers(1, mm)), draw(st.integers(1, mc))
    dt = draw(st.sampled_from(['f4', 'f8']))
    A = draw(csrs(rows, mids, values=True, dtype=dt, **kw))
    B = draw(csrs(mids, cols, values=True, dtype=dt, **kw))
    return A, B


@given(st.data(), csrs(values=True))
def test_mult_vec(kernel, data, csra):
    "tests/test_mult_vec.py:12-24"
    md = csra.to_scipy().toarray()
    v = data.draw(finite_arrays(csra.ncols))
    prod = csra.mult_vec(v)
    assert prod.shape == (csra.nrows,)
    assert prod == approx(md @ v, nan_ok=True, rel=1.0e-5, abs=1.0e-10)
    # and the oracle, at the north-star tolerance
    ref = orc.mult_vec(csra, v)
    assert_values_close(prod, ref, 1e-5 if csra.values.dtype == np.float32 else 1e-10, spmv_terms(csra, v))


@given(st.data(), csrs(values=False))
def test_mult_vec_novalue(kernel, data, csra):
    "tests/test_mult_vec.py:27-39"
    v = data.draw(finite_arrays(csra.ncols))
    prod = csra.mult_vec(v)
    assert prod.shape == (csra.nrows,)
    assert prod == approx(csra.to_scipy() @ v, nan_ok=True)


@given(st.data(), csrs(values=True), st.sampled_from(['f4', 'i8', 'i4']))
def test_mult_vec_x_dtypes(kernel, data, csra, xdt):
    v = data.draw(finite_arrays(csra.ncols, min_value=-100, max_value=100)).astype(xdt)
    prod = csra.mult_vec(v)
    assert prod.dtype == np.float64
    ref = orc.mult_vec(csra, v)
    f4 = csra.values.dtype == np.float32 or v.dtype == np.float32
    assert_values_close(prod, ref, 1e-5 if f4 else 1e-10, spmv_terms(csra, v))


@given(csrs())
def test_make_handle(kernel, csr):
    "tests/test_handles.py:10-21"
    h = kernel.to_handle(csr)
    try:
        assert h is not None
        c2 = kernel.from_handle(h)
        assert c2.nrows == csr.nrows
        assert c2.ncols == csr.ncols
        assert c2.nnz == csr.nnz
    finally:
        kernel.release_handle(h)


def _check_product(prod, dprod, A, B):
    """The reference's check (dense rows, rel 1e-5 / abs 1e-10, tests/test_multiply.py:40-44) for
    float64 inputs.  The reference only draws float64 pairs; for float32 pairs both the numba
    kernel and this one round every product to float32 (multiply.py:120), so the absolute
    tolerance is scaled by the magnitude of the summed terms."""
    nrows = prod.nrows
    if prod.nnz > 0:
        assert prod.values is not None
        assert np.all(prod.values != 0)
    f4 = A.values.dtype == np.float32 or B.values.dtype == np.float32
    atol = 1e-5 * float(np.abs(dprod).max(initial=0.0)) if f4 else 1.0e-10
    dense = prod.to_scipy().toarray()
    for i in range(nrows):
        assert dense[i, :] == approx(dprod[i, :], rel=1.0e-5, abs=atol)


@given(st.data())
def test_multiply(kernel, data):
    "tests/test_multiply.py:14-44 + kernel-level structure parity with the oracle"
    A, B = data.draw(mm_pairs())
    assume(B.nnz < kernel.max_nnz)
    prod = A.multiply(B)
    assert isinstance(prod, CSR)
    assert prod.nrows == A.nrows and prod.ncols == B.ncols
    dprod = A.to_scipy().toarray().astype('f8') @ B.to_scipy().toarray().astype('f8')
    _check_product(prod, dprod, A, B)
    _kernel_level(kernel, A, B, False)


@given(st.data())
def test_multiply_transpose(kernel, data):
    "tests/test_multiply.py:47-79"
    A, B = data.draw(mm_pairs())
    B = B.transpose()
    prod = A.multiply(B, transpose=True)
    assert isinstance(prod, CSR)
    assert prod.nrows == A.nrows and prod.ncols == B.nrows
    dprod = A.to_scipy().toarray().astype('f8') @ B.to_scipy().toarray().astype('f8').T
    _check_product(prod, dprod, A, B)
    _kernel_level(kernel, A, B, True)


def _kernel_level(kernel, A, B, tr):
    ref = orc.mult_abt(A, B) if tr else orc.mult_ab(A, B)
    ah, bh = kernel.to_handle(A), kernel.to_handle(B)
    try:
        ch = kernel.mult_abt(ah, bh) if tr else kernel.mult_ab(ah, bh)
        got = kernel.from_handle(ch)
        kernel.release_handle(ch)
    finally:
        kernel.release_handle(ah)
        kernel.release_handle(bh)
    rp, ci, vs = canonical(ref)
    assert got.rowptrs.dtype == np.int32
    assert np.array_equal(got.rowptrs, rp)
    assert np.array_equal(got.colinds, ci)
    assert_values_close(got.values, vs, value_tol(A, B), spgemm_terms(A, B, tr))


@given(csrs())
def test_kernel_sort_rows(kernel, csr):
    "tests/test_transform.py:77-87 + bit-exact parity with the reference's stable bubble sort"
    tv = np.ones(csr.ncols)
    x1 = csr.mult_vec(tv)
    ref = orc.sort_rows(csr)
    h = kernel.to_handle(csr)
    kernel.order_columns(h)
    c2 = kernel.from_handle(h)
    kernel.release_handle(h)
    assert all(all(np.diff(c2.row_cs(i)) > 0) for i in range(csr.nrows))
    assert c2.mult_vec(tv) == approx(x1)
    assert np.array_equal(c2.colinds, ref.colinds)
    if csr.values is not None:
        assert np.array_equal(c2.values, ref.values)


@given(csrs())
def test_transpose(kernel, csr):
    "tests/test_transpose.py:49-81: rowptrs == scipy .T.tocsr().indptr exactly; dense equality"
    t = csr.transpose()
    assert (t.nrows, t.ncols, t.nnz) == (csr.ncols, csr.nrows, csr.nnz)
    st_ = csr.to_scipy().T.tocsr()
    assert np.array_equal(t.rowptrs, st_.indptr)
    ref = orc.transpose(csr)
    assert np.array_equal(t.colinds, ref.colinds)
    if csr.values is None:
        assert t.values is None
    else:
        assert t.values.dtype == np.float64
        assert np.array_equal(t.values, ref.values)
    t2 = csr.transpose(False)
    assert t2.values is None and np.array_equal(t2.colinds, ref.colinds)


@given(st.data(), st.integers(10, 400))
def test_sharded_paths(kernel, data, lim):
    "tests/test_mkl.py:29-38,76-91: lower max_nnz to force the row-sharding path"
    A, B = data.draw(mm_pairs(max_shape=(60, 40, 60)))
    full = A.multiply(B)
    v = np.linspace(-1, 1, A.ncols)
    yfull = A.mult_vec(v)
    assume(np.diff(A.rowptrs).max(initial=0) <= lim)
    assume(B.nnz <= lim)          # B is uploaded whole (csr.py:560), as in the reference's own sharding tests
    old = kernel.max_nnz
    kernel.max_nnz = lim
    try:
        sh = A.multiply(B)
        ysh = A.mult_vec(v)
    finally:
        kernel.max_nnz = old
    assert sh.nnz == full.nnz
    assert np.array_equal(sh.rowptrs, full.rowptrs)
    assert np.array_equal(sh.colinds, full.colinds)
    # rows are independent, so only the (non-deterministic) order of shared-memory
    # atomics can differ between the two runs
    nz = canonical(orc.mult_ab(A, B))[2] != 0     # multiply() drops stored zeros (csr.py:555)
    if int(nz.sum()) == full.nnz:
        assert_values_close(sh.values, full.values, 1e-12, spgemm_terms(A, B)[nz])
    else:                                          # an exact cancellation on the device only: relative bound alone
        assert_values_close(sh.values, full.values, 1e-12, np.abs(full.values) + 1e-300)
    assert_values_close(ysh, yfull, 1e-12, spmv_terms(A, v))


def _dense_row(m, i):
    "row i as a dense vector (what the reference's CSR.row gives)"
    r = np.zeros(m.ncols, dtype=m.values.dtype)
    r[m.row_cs(i)] = m.row_vs(i)
    return r


@given(csrs(values=True))
def test_mean_center(kernel, csr):
    "tests/test_transform.py:88-125, on the device (CSR.normalize_rows -> csrk_normalize_rows)"
    backup = csr.copy()
    rel_tol, abs_tol = (1.0e-5, 1.0e-3) if csr.values.dtype == np.dtype('f4') else (1.0e-6, 1.0e-10)
    m2 = csr.normalize_rows('center')
    assert len(m2) == csr.nrows and m2.dtype == csr.values.dtype
    rnnz = csr.row_nnzs()
    for i in range(csr.nrows):
        vs, b_vs, b_row = csr.row_vs(i), backup.row_vs(i), _dense_row(backup, i)
        if rnnz[i] > 0:
            assert m2[i] == approx(np.mean(b_vs), rel=rel_tol, abs=abs_tol)
            assert m2[i] == approx(np.sum(b_row) / rnnz[i], rel=rel_tol, abs=abs_tol)
            assert np.mean(vs) == approx(0.0, rel=rel_tol, abs=abs_tol)
            assert vs + m2[i] == approx(b_row[csr.row_cs(i)], rel=rel_tol, abs=abs_tol)
        else:
            assert m2[i] == 0.0
    # and the oracle, tighter
    rvec, ref = orc.normalize_rows(backup, 'center')
    rt = 1e-5 if csr.values.dtype == np.float32 else 1e-12
    scale = float(np.abs(backup.values).max(initial=0.0))
    assert np.all(np.abs(m2.astype(np.float64) - rvec) <= rt * (np.abs(rvec) + scale))
    assert np.all(np.abs(csr.values.astype(np.float64) - ref.values) <= rt * (np.abs(ref.values) + scale))


@given(csrs(values=True))
def test_unit_norm(kernel, csr):
    "tests/test_transform.py:128-147, on the device"
    backup = csr.copy()
    with np.errstate(all='ignore'):
        m2 = csr.normalize_rows('unit')
    assert len(m2) == csr.nrows and m2.dtype == csr.values.dtype
    for i in range(csr.nrows):
        vs, bvs = csr.row_vs(i), backup.row_vs(i)
        if len(vs) > 0:
            assert m2[i] == approx(np.linalg.norm(bvs))
            if m2[i] > 0:
                assert np.linalg.norm(vs) == approx(1.0)
                assert vs * m2[i] == approx(backup.row_vs(i))
            else:
                assert all(np.isnan(vs))
        else:
            assert m2[i] == 0.0


@given(st.data(), st.integers(1, 60), st.integers(1, 60), st.sampled_from(['f4', 'f8', None]))
def test_from_coo_on_device(kernel, data, nrows, ncols, dtype):
    "csr/csr.py:140-169: any COO order, duplicates allowed; bit-identical to the oracle's stable scatter"
    n = data.draw(st.integers(0, 400))
    rows = data.draw(nph.arrays(np.int32, n, elements=st.integers(0, nrows - 1)))
    cols = data.draw(nph.arrays(np.int32, n, elements=st.integers(0, ncols - 1)))
    vals = None if dtype is None else data.draw(finite_arrays(n, dtype=np.dtype(dtype)))
    ref = orc.from_coo(rows, cols, vals, (nrows, ncols))
    h = kernel.from_coo(rows, cols, vals, (nrows, ncols))
    try:
        got = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    assert got.rowptrs.dtype == ref.rowptrs.dtype and np.array_equal(got.rowptrs, ref.rowptrs)
    assert np.array_equal(got.colinds, ref.colinds)
    if dtype is None:
        assert got.values is None
    else:
        assert got.values.dtype == ref.values.dtype and np.array_equal(got.values, ref.values)
    # the host constructor builds the same matrix
    hc = CSR.from_coo(rows, cols, vals, (nrows, ncols))
    assert np.array_equal(hc.rowptrs, ref.rowptrs) and np.array_equal(hc.colinds, ref.colinds)
