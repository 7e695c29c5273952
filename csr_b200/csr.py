"""
``CSR``: the host-side record the ``cuda`` kernel's callers work with.

The kernel module (``csr_b200.kernels.cuda``) is duck-typed on the six fields of the reference's record
(``nrows, ncols, nnz, rowptrs, colinds, values`` -- csr/_struct.py:10-28, dtype rules csr/csr.py:79-100) and
is driven unchanged by the reference's own ``csr.CSR`` (``tests/test_cuda_dropin.py``).  This class exists
for users of this package alone and for what the reference's callers cannot do on a GPU:

* ``multiply`` / ``mult_vec`` (the two callers of the kernel, csr/csr.py:524-590) upload each operand ONCE
  and, when a matrix exceeds ``K.max_nnz``, cut it into row blocks ON THE DEVICE (``csrk_subset_rows``)
  instead of re-slicing and re-uploading host shards (csr/csr.py:599-621); the zero filter of
  csr/csr.py:555 runs on the device before the copy-out;
* ``keep_resident`` keeps one handle per matrix alive across calls -- the
  reference re-creates the handle on every ``mult_vec`` (csr/csr.py:582), which is free for numba and a
  full H2D upload here;
* ``transpose`` / ``sort_rows`` / ``normalize_rows`` / ``from_coo`` run on the device.

Names and argument meaning follow the reference class so that its tests read the same.
"""

from __future__ import annotations

import weakref
from contextlib import contextmanager

import numpy as np

from .kernels import get_kernel, releasing

_I32_MAX = int(np.iinfo(np.intc).max)


class _ResidentHandles:
    """One live kernel handle per resident matrix, keyed by object identity and checked against the
    identity of the three arrays (re-assigning ``values`` or compacting the matrix makes a new handle).
    Writes INTO the arrays are not seen -- "modifying the matrix is not guaranteed to modify handles
    created from it" (csr/kernels/numba/__init__.py:24-26) -- which is why residency is opt-in."""

    def __init__(self):
        self._live = {}      # id(matrix) -> (kernel, handle, signature)
        self._watched = set()

    @staticmethod
    def _signature(m):
        return (id(m.rowptrs), id(m.colinds), id(m._values), m.nnz)

    def handle(self, m, K):
        key = id(m)
        ent = self._live.get(key)
        if ent is not None and ent[0] is K and ent[2] == self._signature(m) and getattr(ent[1], "H", 1):
            return ent[1]
        self.forget(m)
        h = K.to_handle(m)
        self._live[key] = (K, h, self._signature(m))
        if key not in self._watched:           # one finalizer per object, however often the handle is rebuilt
            self._watched.add(key)
            weakref.finalize(m, self._collect, key)
        return h

    def forget(self, m):
        self._release(id(m))

    def _collect(self, key):
        self._watched.discard(key)
        self._release(key)

    def _release(self, key):
        ent = self._live.pop(key, None)
        if ent is not None:
            try:
                ent[0].release_handle(ent[1])
            except Exception:   # interpreter shutdown
                pass


_resident = _ResidentHandles()


def _row_cuts(rowptrs, limit):
    """Row boundaries ``[0, ..., nrows]`` such that no block holds more than ``limit`` entries, each block
    as long as possible (the blocks the reference's ``_shard_rows`` produces one slice at a time,
    csr/csr.py:599-621, found here by searching the absolute row pointers)."""
    assert limit > 0
    rp = np.asarray(rowptrs)
    last = len(rp) - 1
    cuts = [0]
    while int(rp[last]) - int(rp[cuts[-1]]) > limit:
        b = cuts[-1]
        e = int(np.searchsorted(rp, int(rp[b]) + limit, side="right")) - 1
        if e <= b:
            raise ValueError("row too large to fit in target matrix size")
        cuts.append(e)
    cuts.append(last)
    return cuts


class CSR:
    """Compressed sparse row matrix ``nrows, ncols, nnz, rowptrs, colinds, values`` (values optional);
    constructor and attributes as the reference's (csr/csr.py:46-100)."""

    __slots__ = ("nrows", "ncols", "nnz", "rowptrs", "colinds", "_values", "_resident", "__weakref__")

    def __init__(self, nrows, ncols, nnz, rps, cis, vs, _cast=True):
        for dim in (nrows, ncols):
            assert 0 <= dim <= _I32_MAX
        assert nnz >= 0
        self.nrows, self.ncols, self.nnz = int(nrows), int(ncols), int(nnz)
        if _cast:   # csr/csr.py:88-95: int32 structure unless nnz needs 64-bit row pointers; values keep their dtype
            cis = np.require(cis, np.intc, "C")
            rps = np.require(rps, np.intc if nnz <= _I32_MAX else np.int64, "C")
            vs = None if vs is None else np.require(vs, requirements="C")
        self.rowptrs, self.colinds, self._values = rps, cis, vs
        self._resident = False

    # ------------------------------------------------------------ construction / conversion
    @classmethod
    def empty(cls, nrows, ncols, row_nnzs=None, values=True):
        "All-zero structure with the given row lengths (csr/csr.py:102-138)."
        lens = np.zeros(nrows, np.int64) if row_nnzs is None else np.asarray(row_nnzs, dtype=np.int64)
        assert nrows >= 0 and ncols >= 0 and len(lens) == nrows
        rps = np.concatenate([[0], np.cumsum(lens)])
        nnz = int(rps[-1])
        vdt = None if not values and row_nnzs is not None else (np.float64 if values in (True, False) else values)
        return cls(nrows, ncols, nnz, rps, np.zeros(nnz, np.int32), None if vdt is None else np.zeros(nnz, vdt))

    @classmethod
    def from_coo(cls, rows, cols, vals, shape=None, *, rpdtype=np.intc):
        """COO triples -> CSR keeping the COO order inside each row (csr/csr.py:140-173 ->
        csr/structure.py:11-67).  Host version; ``kernels.cuda.from_coo`` builds a handle on the device."""
        rows, cols = np.asarray(rows), np.asarray(cols)
        nnz = len(rows)
        assert len(cols) == nnz and (vals is None or len(vals) == nnz)
        assert np.min(rows, initial=0) >= 0 and np.min(cols, initial=0) >= 0
        if shape is None:
            shape = (int(np.max(rows)) + 1, int(np.max(cols)) + 1)
        nrows, ncols = shape
        assert np.max(rows, initial=0) < max(nrows, 1) and np.max(cols, initial=0) < max(ncols, 1)
        order = np.argsort(rows, kind="stable")
        rps = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=nrows))])
        return cls(nrows, ncols, nnz, rps, cols[order], None if vals is None else np.asarray(vals)[order])

    @classmethod
    def from_scipy(cls, mat, copy=True):
        "csr/csr.py:175-197"
        mat = mat.tocsr(copy=copy) if mat.format != "csr" else mat
        take = (lambda a: np.array(a, copy=True)) if copy else (lambda a: a)
        return cls(mat.shape[0], mat.shape[1], mat.nnz, take(mat.indptr), take(mat.indices), take(mat.data))

    def to_scipy(self):
        "csr/csr.py:199-214 (a structure-only matrix becomes all ones)"
        import scipy.sparse as sps
        data = np.ones(self.nnz) if self._values is None else self._values
        return sps.csr_matrix((data, self.colinds, self.rowptrs), shape=(self.nrows, self.ncols))

    def copy(self, include_values=True, *, copy_structure=True):
        "csr/csr.py:301-322"
        dup = (lambda a: a.copy()) if copy_structure else (lambda a: a)
        vs = self._values.copy() if (include_values and self._values is not None) else None
        return CSR(self.nrows, self.ncols, self.nnz, dup(self.rowptrs), dup(self.colinds), vs)

    def __reduce__(self):
        return (CSR, (self.nrows, self.ncols, self.nnz, self.rowptrs, self.colinds, self._values, False))

    def __str__(self):
        return "<CSR {}x{} ({} nnz)>".format(self.nrows, self.ncols, self.nnz)

    __repr__ = __str__

    # ------------------------------------------------------------ values
    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, vs):
        "csr/csr.py:230-242; a resident handle holds the old values and is dropped"
        if vs is not None:
            if len(vs) < self.nnz:
                raise ValueError("value array too small")
            vs = np.require(vs[:self.nnz], requirements="C")
        self._values = vs
        _resident.forget(self)

    # ------------------------------------------------------------ rows
    def row_extent(self, row):
        "csr/_rows.py:9-13"
        return self.rowptrs[row], self.rowptrs[row + 1]

    def row_cs(self, row):
        sp, ep = self.row_extent(row)
        return self.colinds[sp:ep]

    def row_vs(self, row):
        sp, ep = self.row_extent(row)
        return np.ones(ep - sp) if self._values is None else self._values[sp:ep]

    def row_nnzs(self):
        return np.diff(self.rowptrs)

    def rowinds(self):
        return np.repeat(np.arange(self.nrows, dtype=np.intc), np.diff(self.rowptrs))

    def subset_rows(self, begin, end):
        "Rows [begin, end) as views on this matrix's storage (csr/structure.py:70-81)."
        st, ed = int(self.rowptrs[begin]), int(self.rowptrs[end])
        vs = None if self._values is None else self._values[st:ed]
        return CSR(end - begin, self.ncols, ed - st, self.rowptrs[begin:end + 1] - st, self.colinds[st:ed], vs)

    def _shard_rows(self, tgt_nnz):
        "Host views of the row blocks of at most ``tgt_nnz`` entries (csr/csr.py:599-621)."
        cuts = _row_cuts(self.rowptrs, tgt_nnz)
        return [self.subset_rows(b, e) for b, e in zip(cuts[:-1], cuts[1:])]

    @classmethod
    def _assemble_shards(cls, shards):
        "Row blocks back into one matrix, row pointers rebased in int64 (csr/csr.py:623-650)."
        base = np.cumsum([0] + [s.nnz for s in shards], dtype=np.int64)
        rps = np.concatenate([np.zeros(1, np.int64)] +
                             [np.asarray(s.rowptrs[1:], dtype=np.int64) + b for s, b in zip(shards, base)])
        cis = np.concatenate([s.colinds for s in shards])
        vs = None if shards[0].values is None else np.concatenate([s.values for s in shards])
        assert int(rps[-1]) == int(base[-1]) == len(cis)
        return cls(sum(s.nrows for s in shards), max(s.ncols for s in shards), int(base[-1]), rps, cis, vs)

    def _filter_zeros(self):
        "Drop stored zeros in place (csr/_struct.py:61-79); ``multiply`` does this on the device instead."
        if self._values is None or self._values.all():
            return
        keep = self._values != 0
        kept_before = np.concatenate([[0], np.cumsum(keep)])
        self.rowptrs[:] = kept_before[self.rowptrs]
        self.colinds, self._values, self.nnz = self.colinds[keep], self._values[keep], int(kept_before[-1])
        _resident.forget(self)

    # ------------------------------------------------------------ device residency
    def keep_resident(self, flag=True):
        """Keep this matrix's kernel handle alive between ``mult_vec`` / ``multiply`` / ``normalize_rows``
        calls (see :class:`_ResidentHandles`).  Returns ``self``."""
        self._resident = bool(flag)
        if not flag:
            _resident.forget(self)
        return self

    @contextmanager
    def _on_device(self, K):
        "The matrix as a kernel handle: the resident one, or a temporary released on exit."
        if self._resident:
            yield _resident.handle(self, K)
        else:
            with releasing(K.to_handle(self), K) as h:
                yield h

    @contextmanager
    def _row_blocks(self, K):
        """Handles of the row blocks of at most ``K.max_nnz`` entries: the whole matrix when it fits (the
        common case); else one upload per block (csr/csr.py:558-566,581-590: no handle may exceed ``max_nnz``)."""
        if self.nnz <= K.max_nnz:
            with self._on_device(K) as h:
                yield [h]
            return
        blocks = []
        try:
            for shard in self._shard_rows(K.max_nnz):
                blocks.append(K.to_handle(shard))
            yield blocks
        finally:
            for blk in blocks:
                K.release_handle(blk)

    # ------------------------------------------------------------ the kernel's callers
    def multiply(self, other, transpose=False):
        """``self @ other`` or ``self @ other.T`` (csr/csr.py:524-567): both operands uploaded once, one
        product per row block, stored zeros dropped on the device (csr/csr.py:555), blocks copied out
        and concatenated."""
        assert self.ncols == (other.ncols if transpose else other.nrows)
        K = get_kernel()
        product = K.mult_abt if transpose else K.mult_ab
        parts = []
        with other._on_device(K) as b_h, self._row_blocks(K) as blocks:
            for a_h in blocks:
                with releasing(product(a_h, b_h), K) as c_h:
                    K.filter_zeros(c_h)
                    parts.append(K.from_handle(c_h))
        return parts[0] if len(parts) == 1 else CSR._assemble_shards(parts)

    def mult_vec(self, v):
        "``self @ v`` as a float64 vector (csr/csr.py:569-590)."
        v = np.asarray(v)
        assert v.shape == (self.ncols,)
        K = get_kernel()
        with self._row_blocks(K) as blocks:
            ys = [K.mult_vec(h, v) for h in blocks]
        return ys[0] if len(ys) == 1 else np.concatenate(ys)

    def normalize_rows(self, normalization):
        """Normalise the rows in place on the device and return the per-row means (``'center'``) or norms
        (``'unit'``) (csr/csr.py:443-469 -> csr/transform.py:13-66); an all-zero row under ``'unit'``
        becomes NaN, as in the reference."""
        if normalization not in ("center", "unit"):
            raise ValueError("unknown normalization: " + normalization)
        if self._values is None:
            raise ValueError("normalize_rows needs a matrix with values")
        K = get_kernel()
        vs = np.ascontiguousarray(self._values)
        vecs, at = [], 0
        with self._row_blocks(K) as blocks:   # rows are independent: block by block
            for h in blocks:
                vecs.append(K.normalize_rows(h, normalization, values_out=vs[at:at + h.nnz]))
                at += h.nnz
        if vs is not self._values:
            self._values[...] = vs
        return vecs[0] if len(vecs) == 1 else np.concatenate(vecs)

    def transpose(self, include_values=True):
        "The kernel's stable device transpose (csr/structure.py:172-247); values come back float64."
        K = get_kernel()
        with self._on_device(K) as h, releasing(K.transpose(h, include_values), K) as t:
            return K.from_handle(t)

    def transpose_structure(self):
        return self.transpose(False)

    def sort_rows(self):
        "Sort every row by column in place through the kernel's ``order_columns`` (csr/structure.py:156-169)."
        K = get_kernel()
        with releasing(K.to_handle(self), K) as h:
            K.order_columns(h)
            s = K.from_handle(h)
        self.colinds[:] = s.colinds
        if self._values is not None:
            self._values[:] = s.values
        _resident.forget(self)
