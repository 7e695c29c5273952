"""
Generate tests/golden/golden.npz by running the UNMODIFIED reference
(lenskit/csr v0.5.2, numba kernel) imported from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every case is seeded; the file stores inputs AND the reference's outputs, exactly
as the numba kernel returned them (kernel level, before ``_filter_zeros``; SpGEMM
columns in the reference's reverse-first-touch order).
"""

import os
import sys

os.environ["CSR_KERNEL"] = "numba"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

from csr import CSR  # noqa: E402
from csr.kernels import get_kernel  # noqa: E402

K = get_kernel("numba")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
store = {}
index = []


def put(prefix, m):
    store[f"{prefix}.shape"] = np.array([m.nrows, m.ncols, m.nnz], np.int64)
    store[f"{prefix}.rowptrs"] = np.array(m.rowptrs)
    store[f"{prefix}.colinds"] = np.array(m.colinds)
    if m.values is not None:
        store[f"{prefix}.values"] = np.array(m.values)


def rand_csr(rng, nrows, ncols, density, dtype="f8", values=True, sort=False, dup=0,
             signed=True, ints=False, powerlaw=False, rp64=False):
    """COO in random order -> CSR.from_coo (keeps COO order inside rows)."""
    nnz = int(round(nrows * ncols * density))
    if powerlaw:
        w = 1.0 / np.arange(1, nrows + 1) ** 0.9
        w = rng.permutation(w / w.sum())
        rows = rng.choice(nrows, nnz, p=w).astype(np.int32)
        cw = 1.0 / np.arange(1, ncols + 1) ** 0.8
        cols = rng.choice(ncols, nnz, p=cw / cw.sum()).astype(np.int32)
        coords = np.unique(rows.astype(np.int64) * ncols + cols)
        coords = rng.permutation(coords)
        rows, cols = (coords // ncols).astype(np.int32), (coords % ncols).astype(np.int32)
    else:
        coords = rng.choice(nrows * ncols, nnz, replace=False) if nnz else np.zeros(0, np.int64)
        rows, cols = (coords % nrows).astype(np.int32), (coords // nrows).astype(np.int32)
    if dup and len(rows):
        pick = rng.integers(0, len(rows), dup)
        rows = np.concatenate([rows, rows[pick]])
        cols = np.concatenate([cols, cols[pick]])
    n = len(rows)
    vals = None
    if values:
        if ints:  # small integers: exact cancellation is likely in products
            vals = rng.integers(-2, 3, n).astype(dtype)
            vals[vals == 0] = 1
        elif signed:
            vals = rng.normal(size=n).astype(dtype)
        else:
            vals = rng.uniform(0.5, 5.0, n).astype(dtype)
    m = CSR.from_coo(rows, cols, vals, (nrows, ncols))
    if sort:
        m.sort_rows()
    if rp64:
        m = CSR(m.nrows, m.ncols, m.nnz, m.rowptrs.astype(np.int64), m.colinds, m.values, _cast=False)
    return m


def case_mult_vec(name, m, x):
    h = K.to_handle(m)
    y = K.mult_vec(h, x)
    put(f"{name}.a", m)
    store[f"{name}.x"] = x
    store[f"{name}.y"] = y
    index.append(("mult_vec", name))


def case_mult(name, a, b, transpose):
    c = K.mult_abt(a, b) if transpose else K.mult_ab(a, b)
    put(f"{name}.a", a)
    put(f"{name}.b", b)
    put(f"{name}.c", c)
    # what CSR.multiply hands back to the user (csr.py:555 filters stored zeros)
    f = a.multiply(b, transpose=transpose)
    put(f"{name}.cf", f)
    index.append(("mult_abt" if transpose else "mult_ab", name))


def case_transpose(name, m):
    put(f"{name}.a", m)
    put(f"{name}.t", m.transpose())
    put(f"{name}.ts", m.transpose(False))
    index.append(("transpose", name))


def case_sort(name, m):
    put(f"{name}.a", m)
    s = CSR(m.nrows, m.ncols, m.nnz, m.rowptrs.copy(), m.colinds.copy(),
            None if m.values is None else m.values.copy(), _cast=False)  # keep the rowptr dtype
    h = K.to_handle(s)
    K.order_columns(h)
    put(f"{name}.s", K.from_handle(h))
    index.append(("sort_rows", name))


rng = np.random.default_rng(20261017)

# ---- mult_vec (numba/__init__.py:55-67)
specs = [
    ("mv_f8_small", dict(nrows=7, ncols=5, density=0.4), "f8"),
    ("mv_f8_empty", dict(nrows=4, ncols=6, density=0.0), "f8"),
    ("mv_f8_1x1", dict(nrows=1, ncols=1, density=1.0), "f8"),
    ("mv_f4_x4", dict(nrows=60, ncols=45, density=0.2, dtype="f4"), "f4"),
    ("mv_f4_x8", dict(nrows=33, ncols=70, density=0.3, dtype="f4"), "f8"),
    ("mv_f8_x4", dict(nrows=33, ncols=70, density=0.3, dtype="f8"), "f4"),
    ("mv_nv_x8", dict(nrows=50, ncols=40, density=0.25, values=False), "f8"),
    ("mv_nv_x4", dict(nrows=50, ncols=40, density=0.25, values=False), "f4"),
    ("mv_dup", dict(nrows=20, ncols=15, density=0.3, dup=25), "f8"),
    ("mv_rp64", dict(nrows=40, ncols=40, density=0.2, rp64=True), "f8"),
    ("mv_powerlaw", dict(nrows=400, ncols=300, density=0.08, powerlaw=True, signed=False), "f8"),
    ("mv_powerlaw_f4", dict(nrows=500, ncols=2000, density=0.02, powerlaw=True, dtype="f4"), "f4"),
    ("mv_wide_row", dict(nrows=3, ncols=5000, density=0.6), "f8"),
    ("mv_tall", dict(nrows=3000, ncols=4, density=0.3), "f8"),
]
for name, kw, xdt in specs:
    m = rand_csr(rng, **kw)
    x = rng.normal(size=m.ncols).astype(xdt)
    case_mult_vec(name, m, x)
# integer x is accepted by the numba kernel; result is float64
m = rand_csr(rng, 12, 9, 0.5)
case_mult_vec("mv_int_x", m, rng.integers(-5, 6, m.ncols).astype(np.int64))

# ---- mult_ab / mult_abt (multiply.py:13-57)
mm = [
    ("ab_f8", (17, 23, 11), 0.2, "f8", "f8", {}),
    ("ab_f4", (30, 25, 40), 0.15, "f4", "f4", {}),
    ("ab_f4_f8", (12, 19, 21), 0.25, "f4", "f8", {}),
    ("ab_f8_f4", (12, 19, 21), 0.25, "f8", "f4", {}),
    ("ab_cancel", (25, 20, 25), 0.3, "f8", "f8", dict(ints=True)),
    ("ab_sorted", (40, 30, 35), 0.2, "f8", "f8", dict(sort=True)),
    ("ab_dup", (15, 12, 14), 0.3, "f8", "f8", dict(dup=10)),
    ("ab_empty_a", (6, 8, 5), 0.0, "f8", "f8", {}),
    ("ab_1x1", (1, 1, 1), 1.0, "f8", "f8", {}),
    ("ab_dense", (20, 20, 20), 1.0, "f8", "f8", {}),
    ("ab_powerlaw", (300, 250, 280), 0.05, "f8", "f8", dict(powerlaw=True, signed=False)),
    ("ab_wide", (5, 40, 3000), 0.3, "f8", "f8", {}),
]
for name, (r, k, c), dens, adt, bdt, kw in mm:
    a = rand_csr(rng, r, k, dens, dtype=adt, **kw)
    b = rand_csr(rng, k, c, dens if name != "ab_empty_a" else 0.3, dtype=bdt, **kw)
    case_mult(name, a, b, False)
    bt = rand_csr(rng, c, k, dens if name != "ab_empty_a" else 0.3, dtype=bdt, **kw)
    case_mult(name.replace("ab_", "abt_"), a, bt, True)
# rp64 inputs
a = rand_csr(rng, 14, 16, 0.3, rp64=True)
b = rand_csr(rng, 16, 13, 0.3, rp64=True)
case_mult("ab_rp64", a, b, False)
# item-item shape: M (items x users) times itself transposed
mat = rand_csr(rng, 120, 400, 0.05, powerlaw=True, signed=False, sort=True)
case_mult("abt_itemitem", mat, mat, True)

# ---- transpose (structure.py:172-247)
# known-answer vector from tests/test_transpose.py:11-27
ka = CSR.from_coo(np.array([0, 0, 1, 3]), np.array([1, 2, 0, 1]), np.array([0, 1, 2, 3], dtype=np.float64), (4, 3))
case_transpose("tr_known", ka)
for name, kw in [
    ("tr_f8", dict(nrows=30, ncols=22, density=0.2)),
    ("tr_f4", dict(nrows=25, ncols=40, density=0.2, dtype="f4")),
    ("tr_nv", dict(nrows=25, ncols=40, density=0.2, values=False)),
    ("tr_rp64", dict(nrows=19, ncols=31, density=0.3, rp64=True)),
    ("tr_empty", dict(nrows=5, ncols=7, density=0.0)),
    ("tr_dup", dict(nrows=15, ncols=10, density=0.3, dup=20)),
    ("tr_powerlaw", dict(nrows=500, ncols=300, density=0.05, powerlaw=True)),
    ("tr_tall", dict(nrows=2000, ncols=3, density=0.4)),
    ("tr_wide", dict(nrows=3, ncols=2000, density=0.4)),
]:
    case_transpose(name, rand_csr(rng, **kw))

# ---- order_columns -> sort_rows (structure.py:156-169)
for name, kw in [
    ("so_f8", dict(nrows=30, ncols=50, density=0.3)),
    ("so_f4", dict(nrows=30, ncols=50, density=0.3, dtype="f4")),
    ("so_nv", dict(nrows=30, ncols=50, density=0.3, values=False)),
    ("so_dup", dict(nrows=10, ncols=12, density=0.4, dup=30)),
    ("so_long", dict(nrows=4, ncols=3000, density=0.5)),
    ("so_rp64", dict(nrows=20, ncols=30, density=0.3, rp64=True)),
]:
    case_sort(name, rand_csr(rng, **kw))

store["__index__"] = np.array([f"{k}:{n}" for k, n in index])
np.savez_compressed(OUT, **store)
print(f"wrote {OUT}: {len(index)} cases, {os.path.getsize(OUT) / 1024:.0f} KiB")
