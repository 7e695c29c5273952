"""Shared helpers for the parity tests."""

import numpy as np

from csr_b200 import CSR


def cases(z, kind):
    return [s.split(":", 1)[1] for s in z["__index__"] if s.startswith(kind + ":")]


def gmat(z, prefix, cls=CSR):
    "Rebuild a matrix stored by make_golden.put(); keeps the stored dtypes exactly."
    nrows, ncols, nnz = (int(v) for v in z[f"{prefix}.shape"])
    vs = z[f"{prefix}.values"] if f"{prefix}.values" in z.files else None
    return cls(nrows, ncols, nnz, z[f"{prefix}.rowptrs"], z[f"{prefix}.colinds"], vs, _cast=False)


def canonical(m):
    "Per-row stable sort by column (SURVEY 8c step 3): (rowptrs, colinds, values)."
    rp = np.asarray(m.rowptrs).astype(np.int64)
    rows = np.repeat(np.arange(m.nrows, dtype=np.int64), np.diff(rp))
    order = np.lexsort((np.asarray(m.colinds), rows))
    vs = None if m.values is None else np.asarray(m.values)[order]
    return np.asarray(m.rowptrs), np.asarray(m.colinds)[order], vs


def assert_same_structure(got, ref_rp, ref_ci):
    assert np.asarray(got.rowptrs).dtype == ref_rp.dtype
    assert np.array_equal(np.asarray(got.rowptrs), ref_rp)
    assert np.asarray(got.colinds).dtype == np.int32
    assert np.array_equal(np.asarray(got.colinds), ref_ci)


def value_tol(*mats):
    "rtol of north_star: 1e-10 for float64 inputs, 1e-5 when any input is float32."
    f4 = any(m.values is not None and m.values.dtype == np.float32 for m in mats)
    return 1e-5 if f4 else 1e-10


def assert_values_close(got, ref, rtol, terms=None):
    """|got - ref| <= rtol * max(|ref|, terms), element by element.  ``terms`` is the magnitude of what was
    summed into each element -- ``spgemm_terms`` / ``spmv_terms`` below: sum_k |a_ik||b_kj| resp.
    sum_i |a_i||x_i|, one oracle call on the absolute values -- the standard bound for a re-ordered sum, so
    that cancellation does not turn a different summation order into a false failure.  Without ``terms``
    the bound is purely relative."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    mag = np.abs(ref) if terms is None else np.maximum(np.abs(ref), np.asarray(terms, np.float64))
    err = np.abs(got - ref)
    bad = ~(err <= rtol * mag + 1e-300) & ~(np.isnan(got) & np.isnan(ref)) & ~((got == ref) & np.isinf(ref))
    assert not bad.any(), f"max err {np.nanmax(np.where(bad, err, 0))} at {int(np.argmax(bad))} (rtol {rtol})"


def _abs_mat(m):
    from oracle import oracle as orc
    vs = np.ones(m.nnz) if m.values is None else np.abs(np.asarray(m.values, dtype=np.float64))
    return orc.Mat(m.nrows, m.ncols, m.nnz, np.asarray(m.rowptrs), np.asarray(m.colinds), vs)


def spgemm_terms(a, b, transpose=False):
    "sum_k |a_ik| |b_kj| for every stored element of A B (or A B^T), in canonical (row, column) order."
    from oracle import oracle as orc
    c = (orc.mult_abt if transpose else orc.mult_ab)(_abs_mat(a), _abs_mat(b))
    return canonical(c)[2]


def spmv_terms(a, x):
    "sum_i |a_ri| |x_i| for every row (non-finite x entries count as 0: they are checked separately)."
    from oracle import oracle as orc
    xa = np.abs(np.asarray(x, dtype=np.float64))
    return orc.mult_vec(_abs_mat(a), np.where(np.isfinite(xa), xa, 0.0))
