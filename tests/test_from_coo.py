"""
from_coo (csr/csr.py:140-169 -> csr/structure.py:11-58).

CPU: the oracle restatement against the reference's own outputs (tests/golden/from_coo.npz), bit for bit.
GPU: kernel.from_coo (stable sort by row on the device) against the same vectors, bit for bit, plus a
larger seeded case with many duplicates against the oracle, and the argument checks.
"""

import os

import numpy as np
import pytest

from oracle import oracle as orc

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "from_coo.npz"))
NAMES = [str(n) for n in Z["names"]]


def triples(name):
    nr, nc = (int(v) for v in Z[f"{name}.shape"])
    vals = Z[f"{name}.vals"] if f"{name}.vals" in Z else None
    return Z[f"{name}.rows"], Z[f"{name}.cols"], vals, (nr, nc)


def same_as_reference(name, m):
    assert m.rowptrs.dtype == Z[f"{name}.out_rowptrs"].dtype
    assert np.array_equal(m.rowptrs, Z[f"{name}.out_rowptrs"])
    assert np.array_equal(m.colinds, Z[f"{name}.out_colinds"])
    if f"{name}.out_values" in Z:
        assert m.values.dtype == Z[f"{name}.out_values"].dtype
        assert np.array_equal(m.values, Z[f"{name}.out_values"])
    else:
        assert m.values is None


@pytest.mark.parametrize("name", NAMES)
def test_oracle_from_coo_bit_exact(name):
    same_as_reference(name, orc.from_coo(*triples(name)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_from_coo_bit_exact(kernel, name):
    rows, cols, vals, shape = triples(name)
    h = kernel.from_coo(rows, cols, vals, shape)
    try:
        assert (h.nrows, h.ncols, h.nnz) == (shape[0], shape[1], len(rows))
        same_as_reference(name, kernel.from_handle(h))
    finally:
        kernel.release_handle(h)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f8", "f4", None])
def test_cuda_from_coo_large_with_duplicates(kernel, dtype):
    rng = np.random.default_rng(43)
    nrows, ncols, n = 300000, 1000, 3_000_000           # 19-bit row keys: three 8-bit passes; many duplicates
    rows = rng.integers(0, nrows, n).astype(np.int32)
    rows[rng.random(n) < 0.2] = 12345                    # one very long row
    cols = rng.integers(0, ncols, n).astype(np.int32)
    vals = None if dtype is None else rng.standard_normal(n).astype(dtype)
    ref = orc.from_coo(rows, cols, vals, (nrows, ncols))
    h = kernel.from_coo(rows, cols, vals, (nrows, ncols))
    try:
        got = kernel.from_handle(h)
        x = rng.standard_normal(ncols)
        y = kernel.mult_vec(h, x)                        # the handle is a normal resident matrix
    finally:
        kernel.release_handle(h)
    assert got.rowptrs.dtype == np.int32 and np.array_equal(got.rowptrs, ref.rowptrs)
    assert np.array_equal(got.colinds, ref.colinds)
    assert (got.values is None) if dtype is None else np.array_equal(got.values, ref.values)
    yr = orc.mult_vec(ref, x)
    assert np.allclose(y, yr, rtol=1e-10 if dtype != "f4" else 1e-5, atol=1e-9 * np.abs(yr).max())


@pytest.mark.gpu
def test_cuda_from_coo_argument_checks(kernel):
    with pytest.raises(AssertionError):
        kernel.from_coo([0, 5], [0, 1], None, (3, 3))    # row outside the shape (host assert, csr.py:158)
    with pytest.raises(AssertionError):
        kernel.from_coo([0, 1], [0, -1], None, (3, 3))
    h = kernel.from_coo([2, 0, 2], [1, 1, 0], [1.0, 2.0, 3.0])   # shape inferred
    try:
        m = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    assert (m.nrows, m.ncols) == (3, 2)
    assert m.rowptrs.tolist() == [0, 1, 1, 3] and m.colinds.tolist() == [1, 1, 0] and m.values.tolist() == [2.0, 1.0, 3.0]
