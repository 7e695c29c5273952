"""
Generate tests/golden/from_coo.npz by running the UNMODIFIED reference's CSR.from_coo
(csr/csr.py:140-169 -> csr/structure.py:11-58) imported from /root/reference.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_from_coo.py

Seeded cases: COO triples in random order (duplicates included) and the CSR the reference built.
"""

import os
import sys

os.environ["CSR_KERNEL"] = "numba"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

from csr import CSR  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "from_coo.npz")
store, names = {}, []


def case(name, nrows, ncols, n, dtype, seed, dup=0):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, max(nrows, 1), n).astype(np.int32) if n else np.zeros(0, np.int32)
    cols = rng.integers(0, max(ncols, 1), n).astype(np.int32) if n else np.zeros(0, np.int32)
    if dup and n:
        pick = rng.integers(0, n, dup)
        rows, cols = np.concatenate([rows, rows[pick]]), np.concatenate([cols, cols[pick]])
    vals = None if dtype is None else rng.standard_normal(len(rows)).astype(dtype)
    m = CSR.from_coo(rows, cols, vals, (nrows, ncols))
    store[f"{name}.shape"] = np.array([nrows, ncols], np.int64)
    store[f"{name}.rows"], store[f"{name}.cols"] = rows, cols
    if vals is not None:
        store[f"{name}.vals"] = vals
        store[f"{name}.out_values"] = np.array(m.values)
    store[f"{name}.out_rowptrs"] = np.array(m.rowptrs)
    store[f"{name}.out_colinds"] = np.array(m.colinds)
    names.append(name)


case("c_f8", 50, 40, 600, "f8", 1)
case("c_f4_dups", 30, 20, 500, "f4", 2, dup=200)
case("c_none", 64, 64, 900, None, 3)
case("c_empty_rows", 500, 10, 300, "f8", 4)
case("c_wide_rows", 70000, 50, 20000, "f8", 5)      # 17-bit row keys: the 9-bit digit path
case("c_nothing", 5, 5, 0, "f8", 6)
store["names"] = np.array(names)
np.savez_compressed(OUT, **store)
print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(names), "cases")
