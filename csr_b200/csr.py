"""
Host-side mirror of the reference ``CSR`` class (csr/csr.py:46-692), restated in
plain NumPy for the parts that sit either side of the kernel hot path:

* the six-field record and its dtype rules (csr/csr.py:79-100, csr/_struct.py:10-28);
* the two callers of the kernel, ``multiply`` (csr.py:524-567) and ``mult_vec``
  (csr.py:569-590), with the handle lifecycle, ``max_nnz`` row sharding
  (csr.py:599-650) and the post-multiply zero filter (csr.py:555);
* ``transpose`` / ``sort_rows`` / ``subset_rows`` (csr/structure.py), which here
  run on the device through the active kernel instead of Numba loops;
* construction/conversion helpers the tests need (``from_coo``, ``from_scipy``,
  ``to_scipy``, ``empty``, row accessors, pickling).

It is NOT a Numba structref: the reference's nopython wiring (csr/_wiring.py) is
outside this repo's scope.  The ``cuda`` kernel itself
is duck-typed and accepts the reference's own ``csr.CSR`` objects as well.
"""

from __future__ import annotations

import logging
import weakref

import numpy as np

from .kernels import get_kernel, releasing

INTC = np.iinfo(np.intc)
_log = logging.getLogger(__name__)


class _HandleCache:
    """
    Device residency at the CSR-object level.

    The reference's ``CSR.mult_vec`` makes a kernel handle on every call
    (csr/csr.py:582), which is free for the numba kernel but is a full H2D upload here.
    With ``CSR.keep_resident(True)`` a matrix keeps ONE handle per kernel alive for as
    long as its three arrays are the same objects; the handle is released when the
    matrix is garbage-collected, when its values are re-assigned, or on
    ``CSR.keep_resident(False)``.  In-place writes into the arrays are NOT tracked
    ("modifying the matrix is not guaranteed to modify handles created from it",
    csr/kernels/numba/__init__.py:24-26), so this is opt-in.
    """

    def __init__(self):
        self.entries = {}   # id(csr) -> (kernel, handle, (id(rowptrs), id(colinds), id(values)))

    def get(self, csr, K):
        key = id(csr)
        sig = (id(csr.rowptrs), id(csr.colinds), id(csr._values), csr.nnz)
        ent = self.entries.get(key)
        if ent is not None and ent[0] is K and ent[2] == sig and getattr(ent[1], 'H', 1):
            return ent[1]
        self.drop(csr)
        h = K.to_handle(csr)
        self.entries[key] = (K, h, sig)
        weakref.finalize(csr, self._finalize, key)
        return h

    def drop(self, csr):
        self._finalize(id(csr))

    def _finalize(self, key):
        ent = self.entries.pop(key, None)
        if ent is not None:
            try:
                ent[0].release_handle(ent[1])
            except Exception:  # interpreter shutdown
                pass


_cache = _HandleCache()


class CSR:
    """
    Compressed sparse row matrix: ``nrows, ncols, nnz, rowptrs, colinds, values``
    (values optional).  Same constructor and attribute contract as the reference.
    """

    __slots__ = ("nrows", "ncols", "nnz", "rowptrs", "colinds", "_values", "_resident", "__weakref__")

    def __init__(self, nrows, ncols, nnz, rps, cis, vs, _cast=True):
        # csr.py:79-100
        assert nrows >= 0
        assert nrows <= INTC.max
        assert ncols >= 0
        assert ncols <= INTC.max
        assert nnz >= 0
        self.nrows = int(nrows)
        self.ncols = int(ncols)
        self.nnz = int(nnz)
        if _cast:
            cis = np.require(cis, np.intc, 'C')
            if nnz <= INTC.max:
                rps = np.require(rps, np.intc, 'C')
            else:
                rps = np.require(rps, np.int64, 'C')
            if vs is not None:
                vs = np.require(vs, requirements='C')
        self.rowptrs = rps
        self.colinds = cis
        self._values = vs
        self._resident = False

    def keep_resident(self, flag=True):
        """Keep this matrix's kernel handle alive between ``mult_vec`` / ``multiply`` calls
        (see :class:`_HandleCache`).  Returns ``self``."""
        self._resident = bool(flag)
        if not flag:
            _cache.drop(self)
        return self

    # ------------------------------------------------------------ constructors
    @classmethod
    def empty(cls, nrows, ncols, row_nnzs=None, values=True):
        "csr.py:102-138"
        assert nrows >= 0
        assert ncols >= 0
        if row_nnzs is not None:
            assert len(row_nnzs) == nrows
            nnz = int(np.sum(row_nnzs, dtype=np.int64))
            assert nnz >= 0
            rp_dtype = np.intc if nnz <= INTC.max else np.int64
            rps = np.zeros(nrows + 1, dtype=rp_dtype)
            np.cumsum(row_nnzs, dtype=rp_dtype, out=rps[1:])
            cis = np.zeros(nnz, dtype=np.int32)
            if values is True:
                vs = np.zeros(nnz)
            elif values:
                vs = np.zeros(nnz, dtype=values)
            else:
                vs = None
            return cls(nrows, ncols, nnz, rps, cis, vs)
        # constructors.py:11-23
        return cls(nrows, ncols, 0, np.zeros(nrows + 1, np.intc), np.zeros(0, np.intc), np.zeros(0))

    @classmethod
    def from_coo(cls, rows, cols, vals, shape=None, *, rpdtype=np.intc):
        """csr.py:140-173 + structure.py:11-67: counting sort by row that keeps the
        COO order inside each row (a stable argsort does the same)."""
        rows = np.asarray(rows)
        cols = np.asarray(cols)
        assert np.min(rows, initial=0) >= 0
        assert np.min(cols, initial=0) >= 0
        if shape is not None:
            nrows, ncols = shape
            assert np.max(rows, initial=0) < max(nrows, 1)
            assert np.max(cols, initial=0) < max(ncols, 1)
        else:
            nrows = int(np.max(rows)) + 1
            ncols = int(np.max(cols)) + 1
        nnz = len(rows)
        assert len(cols) == nnz
        assert vals is None or len(vals) == nnz
        order = np.argsort(rows, kind='stable')
        rowptrs = np.zeros(nrows + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=nrows), out=rowptrs[1:])
        out_vals = None if vals is None else np.asarray(vals)[order]
        return cls(nrows, ncols, nnz, rowptrs, cols[order], out_vals)

    @classmethod
    def from_scipy(cls, mat, copy=True):
        "csr.py:175-197"
        import scipy.sparse as sps
        if not sps.isspmatrix_csr(mat):
            mat = mat.tocsr(copy=copy)
        rp = np.require(mat.indptr, np.intc, 'C')
        if copy and rp is mat.indptr:
            rp = rp.copy()
        cs = np.require(mat.indices, np.intc, 'C')
        if copy and cs is mat.indices:
            cs = cs.copy()
        vs = mat.data.copy() if copy else mat.data
        return cls(mat.shape[0], mat.shape[1], mat.nnz, rp, cs, vs)

    def to_scipy(self):
        "csr.py:199-214"
        import scipy.sparse as sps
        values = self.values
        if values is None:
            values = np.full(self.nnz, 1.0)
        return sps.csr_matrix((values, self.colinds, self.rowptrs), shape=(self.nrows, self.ncols))

    # ------------------------------------------------------------------ values
    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, vs):
        "csr.py:230-242"
        if vs is not None:
            if len(vs) < self.nnz:
                raise ValueError('value array too small')
            elif len(vs) > self.nnz:
                vs = vs[:self.nnz]
            vs = np.require(vs, requirements='C')
        self._values = vs
        _cache.drop(self)   # a cached handle holds the old values

    def _required_values(self):
        vs = self.values
        return np.ones(self.nnz) if vs is None else vs

    def _normalize(self, val_dtype=np.float64, ptr_dtype=None):
        "csr.py:264-299"
        if ptr_dtype:
            info = np.iinfo(ptr_dtype)
            if self.nnz > info.max:
                raise ValueError(f'type {ptr_dtype} cannot address {self.nnz} entries')
            rps = np.require(self.rowptrs, ptr_dtype)
        else:
            rps = self.rowptrs
        if val_dtype:
            if self.values is None:
                vs = np.ones(self.nnz, val_dtype)
            else:
                vs = np.require(self.values, val_dtype)
        elif val_dtype is False:
            vs = None
        else:
            vs = self.values
        return CSR(self.nrows, self.ncols, self.nnz, rps, self.colinds, vs, _cast=False)

    def copy(self, include_values=True, *, copy_structure=True):
        "csr.py:301-322"
        values = self.values
        if include_values and values is not None:
            values = np.copy(values)
        else:
            values = None
        rps, cis = self.rowptrs, self.colinds
        if copy_structure:
            rps, cis = np.copy(rps), np.copy(cis)
        return CSR(self.nrows, self.ncols, self.nnz, rps, cis, values)

    # -------------------------------------------------------------- row access
    def row_extent(self, row):
        "csr/_rows.py:9-13"
        return self.rowptrs[row], self.rowptrs[row + 1]

    def row_cs(self, row):
        sp, ep = self.row_extent(row)
        return self.colinds[sp:ep]

    def row_vs(self, row):
        sp, ep = self.row_extent(row)
        if self.values is None:
            return np.full(ep - sp, 1.0)
        return self.values[sp:ep]

    def row(self, row):
        """Dense copy of one row, or a ``k x ncols`` matrix for an array of row indices
        (csr/csr.py:373-387 -> csr/_rows.py:16-87)."""
        row = np.asarray(row, dtype='i4')
        dtype = np.float32 if self.values is None else self.values.dtype
        return self._rows_dense(row, dtype, mask=False)

    def row_mask(self, row):
        "Dense logical array(s) marking the columns stored in the row(s) (csr/csr.py:389-404)."
        return self._rows_dense(np.asarray(row, dtype='i4'), np.bool_, mask=True)

    def _rows_dense(self, row, dtype, mask):
        v = np.zeros(row.shape + (self.ncols,), dtype=dtype)
        if self.nnz == 0:
            return v
        for i, r in enumerate(np.atleast_1d(row)):
            sp, ep = self.row_extent(int(r))
            tgt = v if row.shape == () else v[i, :]
            tgt[self.colinds[sp:ep]] = 1 if (mask or self.values is None) else self.values[sp:ep]
        return v

    def pick_rows(self, rows, *, include_values=True):
        """The given rows (repeats allowed) as a new matrix with int32 rowptrs
        (csr/csr.py:347-364 -> csr/structure.py:85-153)."""
        rows = np.asarray(rows, dtype=np.int64)
        lens = np.diff(self.rowptrs.astype(np.int64))[rows] if len(rows) else np.zeros(0, np.int64)
        rp = np.zeros(len(rows) + 1, dtype=np.int32)
        np.cumsum(lens, out=rp[1:])
        nnz = int(rp[-1])
        # position k of the result comes from rowptrs[rows[r]] + (k - rp[r])
        src = np.repeat(self.rowptrs.astype(np.int64)[rows] - rp[:-1].astype(np.int64), lens) + np.arange(nnz, dtype=np.int64)
        vals = self.values[src] if (include_values and self.values is not None) else None
        return CSR(len(rows), self.ncols, nnz, rp, self.colinds[src].astype(np.int32, copy=False), vals)

    def filter_nnzs(self, filt):
        "Keep the entries where ``filt`` (length nnz) is true (csr/csr.py:494-522)."
        filt = np.asarray(filt)
        if len(filt) != self.nnz:
            raise ValueError('filter has length %d, expected %d' % (len(filt), self.nnz))
        keep = filt.astype(bool)
        rps2 = np.zeros_like(self.rowptrs)
        if self.nnz:
            per_row = np.add.reduceat(np.concatenate([keep.astype(np.int64), [0]]),
                                      np.minimum(self.rowptrs[:-1].astype(np.int64), self.nnz))
            per_row[np.diff(self.rowptrs) == 0] = 0
            np.cumsum(per_row, out=rps2[1:])
        nnz2 = int(rps2[-1])
        assert nnz2 == int(np.sum(keep))
        vs = self.values
        return CSR(self.nrows, self.ncols, nnz2, rps2, self.colinds[keep], None if vs is None else vs[keep])

    def drop_values(self):
        "Remove the value array in place (deprecated in the reference: csr/csr.py:652-661)."
        import warnings
        warnings.warn('drop_values is deprecated', DeprecationWarning)
        self.values = None

    def fill_values(self, value):
        "Set every stored value in place; adds float64 values to a structure-only matrix (csr/csr.py:663-675)."
        if self.values is not None:
            self.values[:] = value
        else:
            self.values = np.full(self.nnz, value, dtype='float64')

    def row_nnzs(self):
        return np.diff(self.rowptrs)

    def rowinds(self):
        "csr/_rows.py:121-128"
        return np.repeat(np.arange(self.nrows, dtype=np.intc), np.diff(self.rowptrs))

    # --------------------------------------------------------------- structure
    def subset_rows(self, begin, end):
        "csr/structure.py:70-81 (views on the parent's storage)."
        st = self.rowptrs[begin]
        ed = self.rowptrs[end]
        rps = self.rowptrs[begin:(end + 1)] - st
        cis = self.colinds[st:ed]
        vs = self.values[st:ed] if self.values is not None else None
        return CSR(end - begin, self.ncols, ed - st, rps, cis, vs)

    def transpose(self, include_values=True):
        """csr/structure.py:172-247, computed by the kernel's stable device
        transpose.  Values come back float64 (structure.py:177)."""
        K = _structure_kernel()
        with releasing(K.to_handle(self), K) as h:
            with releasing(K.transpose(h, include_values), K) as t:
                return K.from_handle(t)

    def transpose_structure(self):
        return self.transpose(False)

    def sort_rows(self):
        "csr/structure.py:156-169 -- in place, through the kernel's order_columns."
        K = _structure_kernel()
        with releasing(K.to_handle(self), K) as h:
            K.order_columns(h)
            s = K.from_handle(h)
        self.colinds[:] = s.colinds
        if self._values is not None:
            self._values[:] = s.values
        _cache.drop(self)

    def _filter_zeros(self):
        """csr/_struct.py:61-79: drop stored zeros in place (host container utility;
        ``multiply`` filters on the device before the copy-out instead)."""
        if self._values is None:
            return
        keep = self._values != 0
        if keep.all():
            return
        pos = np.zeros(self.nnz + 1, np.int64)
        np.cumsum(keep, out=pos[1:])
        self.rowptrs[:] = pos[self.rowptrs]
        self.colinds = self.colinds[keep]
        self._values = self._values[keep]
        self.nnz = int(pos[-1])
        _cache.drop(self)

    # ----------------------------------------------------------------- kernels
    def multiply(self, other, transpose=False):
        """
        ``self @ other`` (or ``self @ other.T``) through the active kernel
        (csr.py:524-567).  ``other`` is uploaded once; ``self`` is row-sharded when
        it exceeds ``K.max_nnz``; stored zeros are dropped from the result.
        """
        if transpose:
            assert self.ncols == other.ncols
        else:
            assert self.ncols == other.nrows

        K = get_kernel()
        dev_filter = getattr(K, 'filter_zeros', None)

        def mul(A, b_h):
            with releasing(K.to_handle(A), K) as a_h:
                if transpose:
                    c_h = K.mult_abt(a_h, b_h)
                else:
                    c_h = K.mult_ab(a_h, b_h)
                with releasing(c_h, K):
                    if dev_filter is not None:
                        dev_filter(c_h)  # csr.py:555, done before the D2H copy
                    crepr = K.from_handle(c_h)
            if dev_filter is None:
                crepr._filter_zeros()
            return crepr

        if self.nnz <= K.max_nnz:
            with releasing(K.to_handle(other), K) as b_h:
                return mul(self, b_h)
        else:
            shards = self._shard_rows(K.max_nnz)
            with releasing(K.to_handle(other), K) as b_h:
                sparts = [mul(s, b_h) for s in shards]
            return CSR._assemble_shards(sparts)

    def mult_vec(self, v):
        "``self @ v`` through the active kernel (csr.py:569-590)."
        v = np.asarray(v)
        assert v.shape == (self.ncols,)
        K = get_kernel()
        if self.nnz <= K.max_nnz:
            if self._resident:
                return K.mult_vec(_cache.get(self, K), v)
            with releasing(K.to_handle(self), K) as h:
                return K.mult_vec(h, v)
        else:
            shards = self._shard_rows(K.max_nnz)
            svs = []
            for s in shards:
                with releasing(K.to_handle(s), K) as h:
                    svs.append(K.mult_vec(h, v))
            return np.concatenate(svs)

    def normalize_rows(self, normalization):
        """Normalise the rows in place and return the per-row means (``'center'``) or norms (``'unit'``)
        (csr.py:443-469 -> transform.py:13-66), computed on the device.  Missing entries are ignored,
        not treated as 0; an all-zero row under ``'unit'`` becomes NaN, as in the reference."""
        if normalization not in ('center', 'unit'):
            raise ValueError('unknown normalization: ' + normalization)
        if self._values is None:
            raise ValueError('normalize_rows needs a matrix with values')
        K = get_kernel()
        vs = np.ascontiguousarray(self._values)
        if self.nnz <= K.max_nnz:
            if self._resident:
                vec = K.normalize_rows(_cache.get(self, K), normalization, values_out=vs)
            else:
                with releasing(K.to_handle(self), K) as h:
                    vec = K.normalize_rows(h, normalization, values_out=vs)
        else:   # rows are independent: normalise shard by shard (csr.py:599-621)
            parts, at = [], 0
            for s in self._shard_rows(K.max_nnz):
                with releasing(K.to_handle(s), K) as h:
                    parts.append(K.normalize_rows(h, normalization, values_out=vs[at:at + s.nnz]))
                at += s.nnz
            vec = np.concatenate(parts)
        if vs is not self._values:
            self._values[...] = vs
        return vec

    def _shard_rows(self, tgt_nnz):
        "csr.py:599-621: split by rows so that every shard has at most tgt_nnz entries."
        assert tgt_nnz > 0
        rest = self
        shards = []
        while rest.nnz > tgt_nnz:
            split = np.searchsorted(rest.rowptrs, tgt_nnz)
            if rest.rowptrs[split] > tgt_nnz:
                if split <= 1:
                    raise ValueError("row too large to fit in target matrix size")
                split -= 1
            _log.debug('splitting %s at %d (rp@s: %d)', rest, split, rest.rowptrs[split])
            shards.append(rest.subset_rows(0, split))
            rest = rest.subset_rows(split, rest.nrows)
        shards.append(rest)
        return shards

    @classmethod
    def _assemble_shards(cls, shards):
        "csr.py:623-650: concatenate row shards, rebasing rowptrs in int64."
        nrows = sum(s.nrows for s in shards)
        ncols = max(s.ncols for s in shards)
        nnz = sum(s.nnz for s in shards)
        rps = np.zeros(nrows + 1, np.int64)
        rs = 0
        for s in shards:
            off = rps[rs]
            re = rs + s.nrows + 1
            rps[rs:re] = s.rowptrs + off
            rs += s.nrows
        assert rps[nrows] == nnz, f'{rps[nrows]} != {nnz}'
        cis = np.concatenate([s.colinds for s in shards])
        assert len(cis) == nnz
        if shards[0].values is not None:
            vs = np.concatenate([s.values for s in shards])
            assert len(vs) == nnz
        else:
            vs = None
        return cls(nrows, ncols, nnz, rps, cis, vs)

    # ------------------------------------------------------------------- misc
    def __str__(self):
        return '<CSR {}x{} ({} nnz)>'.format(self.nrows, self.ncols, self.nnz)

    def __repr__(self):
        return ('<CSR {}x{} ({} nnz) {{\n  rowptrs={}\n  colinds={}\n  values={}\n  dtype={}\n}}>'
                .format(self.nrows, self.ncols, self.nnz, self.rowptrs, self.colinds, self.values,
                        self.values.dtype if self.values is not None else None))

    def __reduce__(self):
        "csr.py:690-692"
        return (CSR, (self.nrows, self.ncols, self.nnz, self.rowptrs, self.colinds, self.values, False))


def _structure_kernel():
    """The kernel used for transpose / sort_rows: the active one if it provides the
    device extras, else the cuda kernel."""
    K = get_kernel()
    if hasattr(K, 'transpose'):
        return K
    return get_kernel('cuda')
