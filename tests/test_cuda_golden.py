"""
GPU: the CUDA path (through the C ABI, via the kernel module) against the outputs
of the unmodified reference stored in tests/golden/golden.npz, and against the
oracle on the same inputs.

Parity bar (BASELINE.md "Parity gate"):
  rowptrs  bit-exact;  colinds bit-exact after canonical per-row sort of the
  reference output;  values rtol 1e-10 (f8 inputs) / 1e-5 (f4 inputs);
  transpose and order_columns bit-exact as-is (structure and values).
"""

import os

import numpy as np
import pytest

from oracle import oracle as orc
from util import (cases, gmat, canonical, assert_same_structure, assert_values_close, value_tol,
                  spgemm_terms, spmv_terms)

pytestmark = pytest.mark.gpu

_Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))


@pytest.mark.parametrize("name", cases(_Z, "mult_vec"))
def test_mult_vec(kernel, golden, name):
    a = gmat(golden, f"{name}.a")
    x = golden[f"{name}.x"]
    ref = golden[f"{name}.y"]
    h = kernel.to_handle(a)
    try:
        y = kernel.mult_vec(h, x)
    finally:
        kernel.release_handle(h)
    assert y.dtype == np.float64 and y.shape == (a.nrows,)
    f4 = (a.values is not None and a.values.dtype == np.float32) or x.dtype == np.float32
    assert_values_close(y, ref, 1e-5 if f4 else 1e-10, spmv_terms(a, x))


@pytest.mark.parametrize("name", cases(_Z, "mult_ab") + cases(_Z, "mult_abt"))
def test_multiply(kernel, golden, name):
    a, b, c = (gmat(golden, f"{name}.{k}") for k in "abc")
    tr = name.startswith("abt_")
    ah, bh = kernel.to_handle(a), kernel.to_handle(b)
    try:
        ch = kernel.mult_abt(ah, bh) if tr else kernel.mult_ab(ah, bh)
        try:
            got = kernel.from_handle(ch)       # kernel level: before _filter_zeros
            stats = kernel.spgemm_stats(ch)
        finally:
            kernel.release_handle(ch)
    finally:
        kernel.release_handle(ah)
        kernel.release_handle(bh)
    assert (got.nrows, got.ncols, got.nnz) == (c.nrows, c.ncols, c.nnz)
    rp, ci, vs = canonical(c)
    assert_same_structure(got, rp, ci)
    assert got.values.dtype == np.float64
    assert_values_close(got.values, vs, value_tol(a, b), spgemm_terms(a, b, tr))
    assert stats["out_nnz"] == c.nnz


@pytest.mark.parametrize("name", cases(_Z, "mult_ab") + cases(_Z, "mult_abt"))
def test_csr_multiply_filters_zeros(kernel, golden, name):
    "CSR.multiply == reference CSR.multiply (stored zeros dropped, csr.py:555)."
    a, b, c, cf = (gmat(golden, f"{name}.{k}") for k in ("a", "b", "c", "cf"))
    got = a.multiply(b, transpose=name.startswith("abt_"))
    if got.nnz != cf.nnz:
        # exact cancellation can differ with summation order only when |value| ~ rounding noise
        pytest.skip("cancellation pattern differs from the reference's summation order")
    rp, ci, vs = canonical(cf)
    assert np.array_equal(got.rowptrs, rp) and np.array_equal(got.colinds, ci)
    keep = canonical(c)[2] != 0   # the bound for the kept entries (cf = c without its stored zeros)
    assert_values_close(got.values, vs, value_tol(a, b), spgemm_terms(a, b, name.startswith("abt_"))[keep])
    assert np.all(got.values != 0)


@pytest.mark.parametrize("name", cases(_Z, "transpose"))
def test_transpose(kernel, golden, name):
    a, t, ts = (gmat(golden, f"{name}.{k}") for k in ("a", "t", "ts"))
    h = kernel.to_handle(a)
    try:
        th = kernel.transpose(h, True)
        got = kernel.from_handle(th)
        kernel.release_handle(th)
        sh = kernel.transpose(h, False)
        gots = kernel.from_handle(sh)
        kernel.release_handle(sh)
    finally:
        kernel.release_handle(h)
    assert (got.nrows, got.ncols, got.nnz) == (t.nrows, t.ncols, t.nnz)
    assert_same_structure(got, t.rowptrs, t.colinds)      # stable order: bit-exact as-is
    if t.values is None:
        assert got.values is None
    else:
        assert got.values.dtype == np.float64
        assert np.array_equal(got.values, t.values)       # pure copy: bit-exact
    assert gots.values is None
    assert_same_structure(gots, ts.rowptrs, ts.colinds)


@pytest.mark.parametrize("name", cases(_Z, "sort_rows"))
def test_order_columns(kernel, golden, name):
    a, s = gmat(golden, f"{name}.a"), gmat(golden, f"{name}.s")
    h = kernel.to_handle(a)
    try:
        kernel.order_columns(h)
        got = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    assert_same_structure(got, s.rowptrs, s.colinds)
    if s.values is None:
        assert got.values is None
    else:
        assert got.values.dtype == s.values.dtype
        assert np.array_equal(got.values, s.values)       # stable among equal columns


@pytest.mark.parametrize("name", cases(_Z, "mult_vec") + cases(_Z, "transpose"))
def test_handle_roundtrip(kernel, golden, name):
    "to_handle -> from_handle preserves every array bit for bit (tests/test_handles.py:10-21)."
    a = gmat(golden, f"{name}.a")
    h = kernel.to_handle(a)
    try:
        b = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    assert (b.nrows, b.ncols, b.nnz) == (a.nrows, a.ncols, a.nnz)
    assert b.rowptrs.dtype == a.rowptrs.dtype and np.array_equal(b.rowptrs, a.rowptrs)
    assert np.array_equal(b.colinds, a.colinds)
    if a.values is None:
        assert b.values is None
    else:
        assert b.values.dtype == a.values.dtype and np.array_equal(b.values, a.values)


def test_filter_zeros_device(kernel, golden):
    c = gmat(golden, "ab_cancel.c")
    assert np.any(c.values == 0)
    ref = orc.filter_zeros(c)
    h = kernel.to_handle(c)
    try:
        kernel.filter_zeros(h)
        got = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    assert got.nnz == ref.nnz
    assert np.array_equal(got.rowptrs, ref.rowptrs)
    assert np.array_equal(got.colinds, ref.colinds)
    assert np.array_equal(got.values, ref.values)


def test_subset_rows_device(kernel, golden):
    a = gmat(golden, "mv_powerlaw.a")
    h = kernel.to_handle(a)
    try:
        for b, e in [(0, 0), (0, a.nrows), (17, 230), (a.nrows - 1, a.nrows)]:
            sh = kernel.subset_rows(h, b, e)
            got = kernel.from_handle(sh)
            kernel.release_handle(sh)
            ref = a.subset_rows(b, e)
            assert (got.nrows, got.ncols, got.nnz) == (ref.nrows, ref.ncols, ref.nnz)
            assert np.array_equal(got.rowptrs, ref.rowptrs)
            assert np.array_equal(got.colinds, ref.colinds)
            assert np.array_equal(got.values, ref.values)
    finally:
        kernel.release_handle(h)
