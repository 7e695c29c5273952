"""
CPU: pin the oracle (oracle/csr_oracle.c) against outputs of the unmodified
reference (tests/golden/golden.npz, written by tests/golden/make_golden.py with the
reference's numba kernel).  Structure AND values must be bit-exact: the oracle
performs the same operations in the same order and types.
"""

import numpy as np
import pytest

from oracle import oracle as orc

from util import cases, gmat

import os
_Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", cases(_Z, "mult_vec"))
def test_mult_vec(golden, name):
    a = gmat(golden, f"{name}.a")
    y = orc.mult_vec(a, golden[f"{name}.x"])
    assert same(y, golden[f"{name}.y"])


@pytest.mark.parametrize("name", cases(_Z, "mult_ab") + cases(_Z, "mult_abt"))
def test_multiply(golden, name):
    a, b, c = (gmat(golden, f"{name}.{k}") for k in "abc")
    got = orc.mult_abt(a, b) if name.startswith("abt_") else orc.mult_ab(a, b)
    assert (got.nrows, got.ncols, got.nnz) == (c.nrows, c.ncols, c.nnz)
    assert same(got.rowptrs, c.rowptrs)
    assert same(got.colinds, c.colinds)      # the reference's own (reverse first-touch) order
    assert same(got.values, c.values)
    # what CSR.multiply returns: the same with stored zeros removed (csr.py:555)
    f = orc.filter_zeros(got)
    cf = gmat(golden, f"{name}.cf")
    assert f.nnz == cf.nnz
    assert same(f.rowptrs, cf.rowptrs) and same(f.colinds, cf.colinds) and same(f.values, cf.values)


@pytest.mark.parametrize("name", cases(_Z, "transpose"))
def test_transpose(golden, name):
    a = gmat(golden, f"{name}.a")
    t, ts = gmat(golden, f"{name}.t"), gmat(golden, f"{name}.ts")
    got = orc.transpose(a)
    assert (got.nrows, got.ncols, got.nnz) == (t.nrows, t.ncols, t.nnz)
    assert same(got.rowptrs, t.rowptrs) and same(got.colinds, t.colinds)
    if t.values is None:
        assert got.values is None
    else:
        assert same(got.values, t.values)
    gs = orc.transpose(a, False)
    assert gs.values is None
    assert same(gs.rowptrs, ts.rowptrs) and same(gs.colinds, ts.colinds)


def test_transpose_known_answer(golden):
    "tests/test_transpose.py:11-27 of the reference: rowptrs of the transpose are [0,1,3,4]."
    a = gmat(golden, "tr_known.a")
    assert np.array_equal(orc.transpose(a).rowptrs, [0, 1, 3, 4])


@pytest.mark.parametrize("name", cases(_Z, "sort_rows"))
def test_sort_rows(golden, name):
    a, s = gmat(golden, f"{name}.a"), gmat(golden, f"{name}.s")
    got = orc.sort_rows(a)
    assert same(got.rowptrs, s.rowptrs) and same(got.colinds, s.colinds)
    if s.values is not None:
        assert same(got.values, s.values)


def test_sym_mm_invariants(golden):
    "tests/test_kernel_numba.py:13-30 of the reference."
    a, b = gmat(golden, "ab_powerlaw.a"), gmat(golden, "ab_powerlaw.b")
    c_rp, c_ci = orc.sym_mm(a, b)
    assert np.all(c_ci >= 0) and np.all(c_ci < b.ncols)
    assert np.all(np.diff(c_rp) >= 0)
    assert len(c_ci) == c_rp[a.nrows]


def test_canonical_equals_sort_rows(golden):
    c = gmat(golden, "ab_powerlaw.c")
    s = orc.sort_rows(c)
    k = orc.canonical(c)
    assert same(s.colinds, k.colinds) and same(s.values, k.values)


def test_threads_match_serial(golden):
    a, b = gmat(golden, "ab_powerlaw.a"), gmat(golden, "ab_powerlaw.b")
    x = np.linspace(-1, 1, a.ncols)
    assert same(orc.mult_vec(a, x), orc.mult_vec_threads(a, x, 3))
    full = orc.mult_ab(a, b)
    parts = orc.mult_threads(a, b, 3)
    assert same(np.concatenate([p.colinds for p in parts]), full.colinds)
    assert same(np.concatenate([p.values for p in parts]), full.values)
