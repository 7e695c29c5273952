"""
ctypes front-end of the CPU oracle (``csr_oracle.c``).

TEST INFRASTRUCTURE ONLY -- parity status: pinned (see ``csr_oracle.c``).  May be
imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, never by ``csr_b200``.

The functions take any object with the six CSR fields of the reference
(``nrows, ncols, nnz, rowptrs, colinds, values`` -- csr/_struct.py:10-28) and
return :class:`Mat` records or plain ndarrays, following the numba kernel's
module contract (csr/kernels/numba/__init__.py, multiply.py).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_ERR = {1: "bad argument / shape mismatch", 2: "out of memory", 3: "out-nnz exceeds INT32_MAX (reference np.intc rowptrs would wrap)"}


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (Makefile next to this file)."""
    src = os.path.join(_HERE, "csr_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "liboracle.so"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64, ip = C.c_void_p, C.c_int32, C.c_int64, C.c_int
        mat = [i32, i32, i64, vp, ip, vp, vp, ip]
        L.orc_mult_vec.argtypes = [i32, i32, i64, vp, ip, vp, vp, ip, vp, ip, vp]
        L.orc_mult_ab.argtypes = mat + mat + [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
        L.orc_mult_abt.argtypes = L.orc_mult_ab.argtypes
        smat = [i32, i32, i64, vp, ip, vp]
        L.orc_sym_mm.argtypes = smat + smat + [vp, C.POINTER(vp), C.POINTER(i64)]
        L.orc_transpose.argtypes = [i32, i32, i64, vp, ip, vp, vp, ip, vp, vp, vp]
        L.orc_sort_rows.argtypes = [i32, vp, ip, vp, vp, ip]
        L.orc_filter_zeros.argtypes = [i32, vp, ip, vp, vp, ip, C.POINTER(i64)]
        L.orc_center_rows.argtypes = [i32, vp, ip, vp, ip, vp]
        L.orc_unit_rows.argtypes = [i32, vp, ip, vp, ip, vp]
        L.orc_from_coo.argtypes = [i32, i64, vp, vp, vp, ip, vp, vp, vp]
        L.orc_free.argtypes = [vp]
        L.orc_free.restype = None
        for f in ("orc_mult_vec", "orc_mult_ab", "orc_mult_abt", "orc_sym_mm", "orc_transpose",
                  "orc_sort_rows", "orc_filter_zeros", "orc_center_rows", "orc_unit_rows", "orc_from_coo"):
            getattr(L, f).restype = C.c_int
        _lib = L
    return _lib


@dataclass
class Mat:
    """Plain CSR record: the six fields of csr/_struct.py:10-28."""
    nrows: int
    ncols: int
    nnz: int
    rowptrs: np.ndarray
    colinds: np.ndarray
    values: Optional[np.ndarray]

    def copy(self) -> "Mat":
        return Mat(self.nrows, self.ncols, self.nnz, self.rowptrs.copy(), self.colinds.copy(),
                   None if self.values is None else self.values.copy())


def as_mat(m) -> Mat:
    """Normalise dtypes the way the CSR constructor does (csr/csr.py:79-100)."""
    rps = np.asarray(m.rowptrs)
    if rps.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
        rps = rps.astype(np.int64)
    rps = np.ascontiguousarray(rps)
    cis = np.ascontiguousarray(np.asarray(m.colinds), dtype=np.int32)
    vs = m.values
    if vs is not None:
        vs = np.ascontiguousarray(np.asarray(vs))
        if vs.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            vs = vs.astype(np.float64)
    return Mat(int(m.nrows), int(m.ncols), int(m.nnz), rps, cis, vs)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _vk(vs):
    return 0 if vs is None else vs.dtype.itemsize


def _args(m: Mat):
    return [m.nrows, m.ncols, m.nnz, _p(m.rowptrs), int(m.rowptrs.dtype.itemsize == 8),
            _p(m.colinds), _p(m.values), _vk(m.values)]


def _check(rc, what):
    if rc:
        if rc == 3:
            raise OverflowError(f"{what}: {_ERR[3]}")
        raise RuntimeError(f"{what}: {_ERR.get(rc, rc)}")


def _x(v):
    v = np.asarray(v)
    if v.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        v = v.astype(np.float64)
    return np.ascontiguousarray(v)


def mult_vec(m, v) -> np.ndarray:
    """numba/__init__.py:55-67"""
    m = as_mat(m)
    v = _x(v)
    assert v.shape == (m.ncols,)
    y = np.empty(m.nrows, np.float64)
    rc = lib().orc_mult_vec(*_args(m), _p(v), v.dtype.itemsize, _p(y))
    _check(rc, "mult_vec")
    return y


def _take(ptr, n, dtype):
    if n == 0:
        out = np.zeros(0, dtype)
    else:
        ct = C.c_int32 if dtype == np.int32 else C.c_double
        out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()
    lib().orc_free(ptr)
    return out


def _mult(fn, a, b, c_ncols) -> Mat:
    a, b = as_mat(a), as_mat(b)
    if a.values is None or b.values is None:
        # the numba kernel fails to type (TypingError) on structure-only inputs
        raise TypeError("mult_ab/mult_abt require value arrays (numba kernel, multiply.py:115,120)")
    c_rp = np.zeros(a.nrows + 1, np.int32)
    ci, vs, nnz = C.c_void_p(), C.c_void_p(), C.c_int64()
    rc = fn(*_args(a), *_args(b), _p(c_rp), C.byref(ci), C.byref(vs), C.byref(nnz))
    _check(rc, fn.__name__)
    n = nnz.value
    return Mat(a.nrows, c_ncols, n, c_rp, _take(ci, n, np.int32), _take(vs, n, np.float64))


def mult_ab(a, b) -> Mat:
    """multiply.py:13-38 (columns in the reference's reverse-first-touch order)."""
    assert a.ncols == b.nrows
    return _mult(lib().orc_mult_ab, a, b, int(b.ncols))


def mult_abt(a, b) -> Mat:
    """multiply.py:41-57"""
    assert a.ncols == b.ncols
    return _mult(lib().orc_mult_abt, a, b, int(b.nrows))


def sym_mm(a, b):
    """multiply.py:60-100 -> (c_rp int32, c_ci int32)"""
    a, b = as_mat(a), as_mat(b)
    assert a.ncols == b.nrows
    c_rp = np.zeros(a.nrows + 1, np.int32)
    ci, nnz = C.c_void_p(), C.c_int64()
    rc = lib().orc_sym_mm(*_args(a)[:6], *_args(b)[:6], _p(c_rp), C.byref(ci), C.byref(nnz))
    _check(rc, "sym_mm")
    return c_rp, _take(ci, nnz.value, np.int32)


def transpose(m, include_values: bool = True) -> Mat:
    """structure.py:172-247"""
    m = as_mat(m)
    with_v = include_values and m.values is not None
    brp = np.zeros(m.ncols + 1, m.rowptrs.dtype)
    bci = np.zeros(m.nnz, np.int32)
    bvs = np.zeros(m.nnz, np.float64) if with_v else None
    rc = lib().orc_transpose(m.nrows, m.ncols, m.nnz, _p(m.rowptrs), int(m.rowptrs.dtype.itemsize == 8),
                             _p(m.colinds), _p(m.values) if with_v else None, _vk(m.values) if with_v else 0,
                             _p(brp), _p(bci), _p(bvs))
    _check(rc, "transpose")
    return Mat(m.ncols, m.nrows, m.nnz, brp, bci, bvs)


def sort_rows(m) -> Mat:
    """structure.py:156-169; returns a sorted COPY (the reference sorts in place)."""
    m = as_mat(m).copy()
    rc = lib().orc_sort_rows(m.nrows, _p(m.rowptrs), int(m.rowptrs.dtype.itemsize == 8),
                             _p(m.colinds), _p(m.values), _vk(m.values))
    _check(rc, "sort_rows")
    return m


def filter_zeros(m) -> Mat:
    """_struct.py:61-79; returns a filtered COPY."""
    m = as_mat(m).copy()
    if m.values is None:
        return m
    nnz = C.c_int64()
    rc = lib().orc_filter_zeros(m.nrows, _p(m.rowptrs), int(m.rowptrs.dtype.itemsize == 8),
                                _p(m.colinds), _p(m.values), _vk(m.values), C.byref(nnz))
    _check(rc, "filter_zeros")
    n = nnz.value
    return Mat(m.nrows, m.ncols, n, m.rowptrs, m.colinds[:n].copy(), m.values[:n].copy())


def normalize_rows(m, normalization: str):
    """csr/csr.py:443-469 -> csr/transform.py:13-66.  Returns (per-row vector, normalised COPY)."""
    m = as_mat(m).copy()
    if m.values is None:
        raise ValueError("normalize_rows needs values")
    vec = np.zeros(m.nrows, m.values.dtype)
    fn = {"center": lib().orc_center_rows, "unit": lib().orc_unit_rows}.get(normalization)
    if fn is None:
        raise ValueError("unknown normalization: " + normalization)
    _check(fn(m.nrows, _p(m.rowptrs), int(m.rowptrs.dtype.itemsize == 8), _p(m.values), _vk(m.values), _p(vec)),
           "normalize_rows")
    return vec, m


def from_coo(rows, cols, vals, shape) -> Mat:
    """csr/csr.py:140-169 -> csr/structure.py:11-58 (shape given; rowptr dtype by the rule of csr.py:90-93)."""
    nrows, ncols = (int(v) for v in shape)
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    nnz = len(rows)
    assert len(cols) == nnz and (vals is None or len(vals) == nnz)
    if vals is not None:
        vals = np.ascontiguousarray(vals)
        if vals.dtype != np.float32:
            vals = vals.astype(np.float64)
    rp = np.zeros(nrows + 1, np.int64)
    oc = np.empty(nnz, np.int32)
    ov = None if vals is None else np.empty_like(vals)
    _check(lib().orc_from_coo(nrows, nnz, _p(rows), _p(cols), _p(vals), _vk(vals), _p(rp), _p(oc), _p(ov)), "from_coo")
    if nnz <= np.iinfo(np.int32).max:
        rp = rp.astype(np.int32)
    return Mat(nrows, ncols, nnz, rp, oc, ov)


def canonical(m: Mat) -> Mat:
    """Canonical per-row column order of a product: a stable
    argsort by (row, col).  Equal to ``sort_rows`` but O(n log n)."""
    m = as_mat(m)
    rows = np.repeat(np.arange(m.nrows, dtype=np.int64), np.diff(m.rowptrs.astype(np.int64)))
    order = np.lexsort((m.colinds, rows))
    return Mat(m.nrows, m.ncols, m.nnz, m.rowptrs.copy(), m.colinds[order],
               None if m.values is None else m.values[order])


# ---- all-cores variants: the reference kernels are nogil (numba/__init__.py:55,
# multiply.py:13,41), so a thread pool over CSR._shard_rows-style row blocks
# (csr/csr.py:599-621) is how a user runs them on every host core without
# touching the reference.  ctypes drops the GIL the same way.

def row_blocks(m: Mat, nblocks: int):
    """Contiguous row blocks of ~equal nnz: split_k = searchsorted(rowptrs, k*nnz/N)."""
    rp = m.rowptrs.astype(np.int64)
    cuts = [0]
    for k in range(1, nblocks):
        s = int(np.searchsorted(rp, (m.nnz * k) // nblocks, side="left"))
        cuts.append(min(max(s, cuts[-1]), m.nrows))
    cuts.append(m.nrows)
    return [(cuts[i], cuts[i + 1]) for i in range(nblocks) if cuts[i + 1] > cuts[i]]


def subset_rows(m: Mat, begin: int, end: int) -> Mat:
    """structure.py:70-81 (views, rebased pointers)."""
    st, ed = int(m.rowptrs[begin]), int(m.rowptrs[end])
    rps = m.rowptrs[begin:end + 1] - m.rowptrs[begin]
    return Mat(end - begin, m.ncols, ed - st, np.ascontiguousarray(rps), m.colinds[st:ed],
               None if m.values is None else m.values[st:ed])


def mult_vec_threads(m, v, threads: int) -> np.ndarray:
    m = as_mat(m)
    v = _x(v)
    blocks = [subset_rows(m, b, e) for b, e in row_blocks(m, threads)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(lambda s: mult_vec(s, v), blocks))
    return np.concatenate(parts) if parts else np.zeros(0)


def mult_threads(a, b, threads: int, transpose_b: bool = False):
    """Row-block-parallel mult_ab / mult_abt; returns the list of block results
    (assembling them is csr.py:623-650 and is not part of the timed kernel)."""
    a, b = as_mat(a), as_mat(b)
    if transpose_b:
        b = transpose(b)
    blocks = [subset_rows(a, s, e) for s, e in row_blocks(a, threads)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda s: mult_ab(s, b), blocks))
