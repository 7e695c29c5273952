"""
GPU: nopython code driving the cuda kernel through csr_b200/kernels/cuda_numba.py:
the same results as the object-mode kernel module, which the other GPU tests pin to the oracle.
"""

import numpy as np
import pytest

numba = pytest.importorskip("numba")
from numba import njit  # noqa: E402

from csr_b200 import synth  # noqa: E402
from csr_b200.kernels import cuda_numba as cn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


@njit
def _power_step(h, x):
    y = cn.mult_vec(h, x)
    return y, np.sqrt((y * y).sum())


@njit
def _gram(h):
    "A A^T from nopython: product, copy-out, release."
    c = cn.mult_abt(h, h)
    nr, nc, nnz, rp, ci, vs = cn.export_arrays(c)
    cn.release_handle(c)
    return nr, nc, nnz, rp, ci, vs


@pytest.mark.parametrize("dtype", ["f8", "f4"])
def test_mult_vec_from_nopython(kernel, dtype):
    A = synth.powerlaw_csr(5000, 3000, 200000, seed=31, dtype=dtype, alpha=1.0)
    h = kernel.to_handle(A)
    try:
        # contiguous floats alias the caller's array; integers and strided views make the nopython
        # wrapper build a TEMPORARY that must stay alive across the native call
        strided = synth.dense_vector(2 * A.ncols, 3, "f8")[::2]
        strided4 = synth.dense_vector(3 * A.ncols, 4, "f4")[1::3]
        for x in (synth.dense_vector(A.ncols, 1, "f8"), synth.dense_vector(A.ncols, 2, "f4"),
                  np.arange(A.ncols, dtype=np.int64) % 7, strided, strided4,
                  (np.arange(A.ncols) % 3 == 0)):
            y, nrm = _power_step(h.H, x)
            ref = kernel.mult_vec(h, x)
            assert y.dtype == np.float64 and np.array_equal(y, ref)
            assert nrm == pytest.approx(np.sqrt((ref * ref).sum()))
            xo = np.ascontiguousarray(x if x.dtype in (np.float32, np.float64) else x.astype(np.float64))
            want = orc.mult_vec(A, xo)
            f4 = dtype == "f4" or xo.dtype == np.float32
            assert np.allclose(y, want, rtol=1e-5 if f4 else 1e-10, atol=(1e-5 if f4 else 1e-10) * np.abs(want).max())
        assert tuple(int(v) for v in cn.dims(h.H)) == (A.nrows, A.ncols, A.nnz, 0, A.values.dtype.itemsize)
        with pytest.raises(ValueError):
            _power_step(h.H, np.zeros(A.ncols + 1))
    finally:
        kernel.release_handle(h)


def test_product_and_export_from_nopython(kernel):
    A = synth.powerlaw_csr(2000, 1500, 60000, seed=37, dtype="f8", alpha=0.8)
    h = kernel.to_handle(A)
    try:
        nr, nc, nnz, rp, ci, vs = _gram(h.H)
        ch = kernel.mult_abt(h, h)
        ref = kernel.from_handle(ch)
        kernel.release_handle(ch)
    finally:
        kernel.release_handle(h)
    assert (nr, nc, nnz) == (ref.nrows, ref.ncols, ref.nnz)
    assert rp.dtype == np.int64 and np.array_equal(rp, ref.rowptrs) and np.array_equal(ci, ref.colinds)
    assert np.allclose(vs, ref.values, rtol=1e-12, atol=0.0)
    want = orc.canonical(orc.mult_abt(A, A))          # and against the oracle, not only CUDA against CUDA
    assert np.array_equal(rp, want.rowptrs) and np.array_equal(ci, want.colinds)
    assert np.allclose(vs, want.values, rtol=1e-10, atol=0.0)


@njit
def _roundtrip(nrows, ncols, nnz, rp, ci, vs, x):
    "to_handle -> mult_vec -> from_handle -> release_handle, all from nopython code."
    h = cn.create(nrows, ncols, nnz, rp, ci, vs)
    y = cn.mult_vec(h, x)
    out = cn.export_arrays(h)
    cn.release_handle(h)
    return y, out


@njit
def _roundtrip_structure(nrows, ncols, nnz, rp, ci, x):
    h = cn.create_structure(nrows, ncols, nnz, rp, ci)
    y = cn.mult_vec(h, x)
    cn.release_handle(h)
    return y


@pytest.mark.parametrize("dtype,rp64", [("f8", False), ("f4", True)])
def test_handle_lifecycle_from_nopython(kernel, dtype, rp64):
    A = synth.powerlaw_csr(3000, 2000, 90000, seed=39, dtype=dtype, alpha=0.9)
    rp = A.rowptrs.astype(np.int64 if rp64 else np.int32)
    x = synth.dense_vector(A.ncols, 5, "f8")
    y, (nr, nc, nnz, rpo, cio, vso) = _roundtrip(A.nrows, A.ncols, A.nnz, rp, A.colinds, A.values, x)
    assert (nr, nc, nnz) == (A.nrows, A.ncols, A.nnz)
    assert np.array_equal(rpo, A.rowptrs) and np.array_equal(cio, A.colinds) and np.array_equal(vso, A.values.astype(np.float64))
    want = orc.mult_vec(A, x)
    tol = 1e-5 if dtype == "f4" else 1e-10
    assert np.allclose(y, want, rtol=tol, atol=tol * np.abs(want).max())
    ys = _roundtrip_structure(A.nrows, A.ncols, A.nnz, rp, A.colinds, x)
    S = orc.Mat(A.nrows, A.ncols, A.nnz, A.rowptrs, A.colinds, None)
    ws = orc.mult_vec(S, x)
    assert np.allclose(ys, ws, rtol=1e-10, atol=1e-10 * np.abs(ws).max())
