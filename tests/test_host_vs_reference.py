"""
CPU, build container only: the host mirror's off-path helpers (row, row_mask, pick_rows, filter_nnzs,
fill_values, rowinds, row_nnzs) against the UNMODIFIED reference imported from /root/reference.
Skipped where the reference is absent (the GPU box); nothing on the GPU path depends on this file.
"""

import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
if not os.path.isdir(os.path.join(REF, "csr")):
    pytest.skip("reference not present", allow_module_level=True)

from csr_b200 import CSR  # noqa: E402


@pytest.fixture(scope="module")
def refmod():
    old_env, old_flag, saved = os.environ.get("CSR_KERNEL"), sys.dont_write_bytecode, sys.path[:]
    os.environ["CSR_KERNEL"] = "numba"       # for the reference's import only; restored below
    sys.dont_write_bytecode = True           # /root/reference is read-only
    sys.path.insert(0, REF)
    try:
        import csr as ref
    finally:
        sys.path[:] = saved
        sys.dont_write_bytecode = old_flag
        if old_env is None:
            os.environ.pop("CSR_KERNEL", None)
        else:
            os.environ["CSR_KERNEL"] = old_env
    yield ref


def pair(ref, seed, values=True, dtype="f8", nrows=40, ncols=30, nnz=300):
    rng = np.random.default_rng(seed)
    coords = rng.choice(nrows * ncols, nnz, replace=False)
    rows, cols = (coords // ncols).astype(np.int32), (coords % ncols).astype(np.int32)
    vals = rng.standard_normal(nnz).astype(dtype) if values else None
    return CSR.from_coo(rows, cols, vals, (nrows, ncols)), ref.CSR.from_coo(rows, cols, vals, (nrows, ncols))


def same(a, b):
    assert (a.nrows, a.ncols, a.nnz) == (b.nrows, b.ncols, b.nnz)
    assert a.rowptrs.dtype == b.rowptrs.dtype and np.array_equal(a.rowptrs, b.rowptrs)
    assert np.array_equal(a.colinds, b.colinds)
    if b.values is None:
        assert a.values is None
    else:
        assert a.values.dtype == b.values.dtype and np.array_equal(a.values, b.values)


@pytest.mark.parametrize("values", [True, False])
def test_rows_masks_and_indices(refmod, values):
    m, r = pair(refmod, 1, values=values, dtype="f4")
    for idx in (0, 7, [3, 3, 9], np.array([], dtype="i4")):
        got, exp = m.row(idx), r.row(idx)
        assert got.dtype == exp.dtype and got.shape == exp.shape and np.array_equal(got, exp)
        gm, em = m.row_mask(idx), r.row_mask(idx)
        assert gm.dtype == em.dtype and gm.shape == em.shape and np.array_equal(gm, em)
    assert np.array_equal(m.rowinds(), r.rowinds()) and np.array_equal(m.row_nnzs(), r.row_nnzs())


@pytest.mark.parametrize("values", [True, False])
def test_pick_rows(refmod, values):
    m, r = pair(refmod, 2, values=values)
    for rows in ([5, 1, 1, 39, 0], [], list(range(40))):
        rows = np.array(rows, dtype=np.int32)
        same(m.pick_rows(rows), r.pick_rows(rows))
        same(m.pick_rows(rows, include_values=False), r.pick_rows(rows, include_values=False))


def test_filter_nnzs_and_fill_values(refmod):
    m, r = pair(refmod, 3)
    filt = m.values > 0.2
    same(m.filter_nnzs(filt), r.filter_nnzs(filt))
    same(m.filter_nnzs(np.zeros(m.nnz, bool)), r.filter_nnzs(np.zeros(r.nnz, bool)))
    with pytest.raises(ValueError):
        m.filter_nnzs(filt[:-1])
    m.fill_values(2.5)
    r.fill_values(2.5)
    same(m, r)
    ms, rs = pair(refmod, 4, values=False)
    ms.fill_values(1.5)
    rs.fill_values(1.5)
    same(ms, rs)
    with pytest.warns(DeprecationWarning):
        ms.drop_values()
    assert ms.values is None
