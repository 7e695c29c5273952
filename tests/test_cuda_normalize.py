"""
GPU: CSR.normalize_rows / kernel.normalize_rows on the device against the reference's
own outputs (tests/golden/normalize.npz) and the oracle.

The device sums in float64 with a lane-strided order, the reference sequentially in the values' dtype, so
values agree to rounding: rtol 1e-12 (float64) / 1e-5 (float32) with an absolute term of the same
relative size times the largest magnitude of the row's inputs (cancellation in the mean).  The reference's
own tests (tests/test_transform.py:94-99) use 1e-6 / 1e-5.
"""

import os

import numpy as np
import pytest

from csr_b200 import CSR, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "normalize.npz"))
NAMES = [str(n) for n in Z["names"]]


def gmat(name):
    nr, nc, nnz = (int(v) for v in Z[f"{name}.shape"])
    return CSR(nr, nc, nnz, Z[f"{name}.rowptrs"].copy(), Z[f"{name}.colinds"].copy(), Z[f"{name}.values"].copy())


def close(got, ref, dtype, row_scale):
    rt = 1e-5 if dtype == np.float32 else 1e-12
    g, r = got.astype(np.float64), ref.astype(np.float64)
    with np.errstate(all="ignore"):
        ok = (np.abs(g - r) <= rt * np.abs(r) + rt * row_scale) | (np.isnan(g) & np.isnan(r))
    return bool(np.all(ok))


def scales(m):
    "largest input magnitude of the row each entry / each row belongs to"
    lens = np.diff(m.rowptrs.astype(np.int64))
    a = np.abs(m.values.astype(np.float64))
    row_max = np.zeros(m.nrows)
    nz = lens > 0
    row_max[nz] = np.maximum.reduceat(a, m.rowptrs[:-1][nz].astype(np.int64)) if m.nnz else 0.0
    return row_max, np.repeat(row_max, lens)


@pytest.mark.parametrize("kind", ["center", "unit"])
@pytest.mark.parametrize("name", NAMES)
def test_kernel_level_against_reference(kernel, name, kind):
    m = gmat(name)
    row_max, ent_max = scales(m)
    h = kernel.to_handle(m)
    try:
        vec = kernel.normalize_rows(h, kind)
        got = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    gv, gvals = Z[f"{name}.{kind}.vec"], Z[f"{name}.{kind}.values"]
    assert vec.dtype == gvals.dtype and got.values.dtype == gvals.dtype
    assert np.array_equal(got.rowptrs, m.rowptrs) and np.array_equal(got.colinds, m.colinds)
    assert close(vec, gv, gvals.dtype, row_max)
    # normalised values: centred ones are differences (absolute scale = row magnitude), unit ones are O(1)
    assert close(got.values, gvals, gvals.dtype, ent_max if kind == "center" else 1.0)


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("kind", ["center", "unit"])
def test_csr_level_in_place(kernel, kind, resident):
    A = synth.powerlaw_csr(20000, 3000, 600000, seed=23, dtype="f8", alpha=0.9)
    A.values[:] = A.values - 2.5            # signed
    B = A.copy()
    if resident:
        B.keep_resident(True)
    vals_obj = B.values
    vec = B.normalize_rows(kind)
    assert B.values is vals_obj             # mutated in place, like the reference
    rvec, ref = orc.normalize_rows(orc.as_mat(A), kind)
    row_max, ent_max = scales(A)
    assert close(vec, rvec, np.float64, row_max)
    assert close(B.values, ref.values, np.float64, ent_max if kind == "center" else 1.0)
    lens = np.diff(A.rowptrs)
    nz = lens > 0
    if kind == "center":      # row means vanish
        sums = np.add.reduceat(B.values, A.rowptrs[:-1][nz].astype(np.int64))
        assert np.all(np.abs(sums) <= 1e-9 * row_max[nz] * lens[nz])
    else:                     # rows are unit vectors
        ss = np.add.reduceat(B.values ** 2, A.rowptrs[:-1][nz].astype(np.int64))
        assert np.allclose(ss, 1.0, rtol=1e-12)
    if resident:              # the resident handle holds the normalised values too
        x = synth.dense_vector(A.ncols, 3, "f8")
        y = B.mult_vec(x)
        assert np.allclose(y, orc.mult_vec(orc.as_mat(B), x), rtol=1e-10, atol=1e-10)
        B.keep_resident(False)


def test_float32_long_rows_and_errors(kernel):
    A = synth.powerlaw_csr(300, 200000, 2_000_000, seed=29, dtype="f4", alpha=1.0)   # rows up to ~10^5 entries
    for kind in ("center", "unit"):
        B = A.copy()
        vec = B.normalize_rows(kind)
        rvec, ref = orc.normalize_rows(orc.as_mat(A), kind)
        assert vec.dtype == np.float32 and B.values.dtype == np.float32
        row_max, ent_max = scales(A)
        # the reference's float32 running sum over 10^5 entries is itself only good to ~1e-3 relative
        tol = 3e-3 if kind == "center" else 1e-5
        assert np.all(np.abs(vec.astype(np.float64) - rvec) <= tol * (np.abs(rvec) + row_max))
        assert np.all(np.abs(B.values.astype(np.float64) - ref.values) <= tol * (np.abs(ref.values) + (ent_max if kind == "center" else 1.0)))
    with pytest.raises(ValueError):
        A.normalize_rows("l1")
    with pytest.raises(ValueError):
        CSR(A.nrows, A.ncols, A.nnz, A.rowptrs, A.colinds, None).normalize_rows("center")
    h = kernel.to_handle(CSR(A.nrows, A.ncols, A.nnz, A.rowptrs, A.colinds, None))
    try:
        with pytest.raises(ValueError):
            kernel.normalize_rows(h, "unit")
    finally:
        kernel.release_handle(h)
