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


def assert_values_close(got, ref, rtol, scale=None):
    """|got-ref| <= rtol * (|ref| + scale): `scale` is the magnitude of the terms that
    were summed, so cancellation does not turn a reordering into a false failure."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    if scale is None:
        scale = np.abs(ref).max(initial=0.0)
    err = np.abs(got - ref)
    bound = rtol * (np.abs(ref) + scale) + 1e-300
    bad = ~(err <= bound) & ~(np.isnan(got) & np.isnan(ref))
    assert not bad.any(), f"max err {err.max()} at {int(np.argmax(err))} (rtol {rtol})"


def abs_product_scale(a, b, transpose=False):
    "Row-wise upper bound sum |a||b| for an SpGEMM result, as a dense-free scalar per matrix."
    av = np.abs(a.values).max(initial=0.0) if a.values is not None else 1.0
    bv = np.abs(b.values).max(initial=0.0) if b.values is not None else 1.0
    la = np.diff(np.asarray(a.rowptrs).astype(np.int64)).max(initial=0)
    return float(av) * float(bv) * max(int(la), 1)
