"""
Generate tests/golden/normalize.npz by running the UNMODIFIED reference
(lenskit/csr v0.5.2, csr/transform.py:13-66 through CSR.normalize_rows, csr/csr.py:443-469)
imported from /root/reference.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_normalize.py

Seeded cases; the file stores the inputs and, for both normalisations, the returned
per-row vector and the values the reference left in the matrix.
"""

import os
import sys

os.environ["CSR_KERNEL"] = "numba"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

from csr import CSR  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "normalize.npz")
store = {}
names = []


def case(name, nrows, ncols, nnz, dtype, seed, scale=1.0, rp64=False, tweak=None):
    rng = np.random.default_rng(seed)
    coords = rng.choice(nrows * ncols, nnz, replace=False) if nnz else np.zeros(0, np.int64)
    rows, cols = (coords // ncols).astype(np.int32), (coords % ncols).astype(np.int32)
    vals = (rng.standard_normal(nnz) * scale).astype(dtype)
    if tweak is not None:
        tweak(rows, vals)
    m = CSR.from_coo(rows, cols, vals, (nrows, ncols))
    if rp64:
        m = CSR(m.nrows, m.ncols, m.nnz, m.rowptrs.astype(np.int64), m.colinds, m.values)
    store[f"{name}.shape"] = np.array([m.nrows, m.ncols, m.nnz], np.int64)
    store[f"{name}.rowptrs"] = np.array(m.rowptrs)
    store[f"{name}.colinds"] = np.array(m.colinds)
    store[f"{name}.values"] = np.array(m.values)
    for kind in ("center", "unit"):
        c = m.copy()
        with np.errstate(all="ignore"):
            vec = c.normalize_rows(kind)
        store[f"{name}.{kind}.vec"] = np.array(vec)
        store[f"{name}.{kind}.values"] = np.array(c.values)
    names.append(name)


def zero_row(rows, vals):
    vals[rows == 3] = 0.0       # a stored-zero row: unit norm 0 -> NaN values


def tiny_row(rows, vals):
    vals[rows == 5] *= 1e-30    # all-tiny row: the pre-normalisation matters


case("n_f8_small", 40, 30, 300, "f8", 1)
case("n_f4_small", 40, 30, 300, "f4", 2)
case("n_f8_empty_rows", 200, 50, 400, "f8", 3)
case("n_f8_zero_row", 30, 40, 500, "f8", 4, tweak=zero_row)
case("n_f4_tiny", 30, 40, 500, "f4", 5, scale=1e-3, tweak=tiny_row)
case("n_f8_tiny", 30, 40, 500, "f8", 6, scale=1e-200)
case("n_f8_huge", 30, 40, 500, "f8", 7, scale=1e150)
case("n_f8_rp64", 300, 200, 9000, "f8", 8, rp64=True)
case("n_f8_wide", 64, 5000, 40000, "f8", 9)
case("n_f4_wide", 64, 5000, 40000, "f4", 10, scale=3.0)
case("n_f8_nothing", 10, 10, 0, "f8", 11)
store["names"] = np.array(names)
np.savez_compressed(OUT, **store)
print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(names), "cases")
