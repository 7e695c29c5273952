"""
``cuda`` kernel: the B200 (sm_100a) implementation of lenskit/csr's kernel
contract (docs/kernels.rst:61-104; csr/kernel.py:9-16), a drop-in peer of
``csr.kernels.numba`` / ``csr.kernels.mkl``.

Module attributes required by the contract::

    max_nnz, to_handle, from_handle, release_handle, order_columns,
    mult_ab, mult_abt, mult_vec

A handle keeps ``rowptrs / colinds / values`` resident in HBM for its lifetime;
every operation is one call into ``libcsr_cuda.so`` (C ABI: ``include/csrk.h``)
through ctypes, which drops the GIL for the duration (the numba kernels are
``nogil`` too: csr/kernels/numba/__init__.py:55, multiply.py:13,41).
There is no CPU fallback: without the library the import fails, without a B200
the first ``to_handle`` raises.

Differences from the numba kernel, all within the written contract:

* product handles hold columns sorted ascending inside each row (the numba
  kernel emits reverse first-touch order, multiply.py:79-82,94-97);
* ``mult_ab`` / ``mult_abt`` accept structure-only operands (values = 1), as the
  MKL kernel does (csr/kernels/mkl/handle.py:69); numba fails to type them;
* extras beyond the contract, used by the host layer when present:
  ``transpose``, ``filter_zeros``, ``subset_rows``, ``spgemm_stats`` and the
  device-pointer entry ``mult_vec_dev``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _native as N

__all__ = ["max_nnz", "to_handle", "from_handle", "release_handle", "order_columns",
           "mult_ab", "mult_abt", "mult_vec"]

# docs/kernels.rst:98-104.  HBM (180 GB) bounds real inputs long before this; the
# host layer shards rows when nnz exceeds it (csr/csr.py:558,581), and tests lower
# it to force that path (tests/test_mkl.py:29-38).
max_nnz = np.iinfo("i8").max

_INTC_MAX = np.iinfo(np.intc).max

# fail at import, loudly, if the native library is absent (no CPU fallback)
N.lib()


class cuda_h:
    """Opaque handle: an owning reference to a device-resident matrix.

    Mirrors ``mkl_h`` (csr/kernels/mkl/handle.py:24-43): ``H`` is the native
    handle (0 once released), ``nrows/ncols`` are cached on the host.
    """

    __slots__ = ("H", "nrows", "ncols", "nnz", "csr_cls", "__weakref__")

    def __init__(self, H, nrows, ncols, nnz, csr_cls):
        self.H = H
        self.nrows = int(nrows)
        self.ncols = int(ncols)
        self.nnz = int(nnz)
        self.csr_cls = csr_cls

    def __repr__(self):
        state = "released" if not self.H else hex(self.H)
        return f"<cuda_h {self.nrows}x{self.ncols} ({self.nnz} nnz) {state}>"


def _default_csr_cls():
    from ..csr import CSR
    return CSR


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _live(h: cuda_h):
    if not isinstance(h, cuda_h):
        raise TypeError(f"expected a cuda kernel handle, got {type(h).__name__}")
    if not h.H:
        raise ValueError("handle has been released")
    return C.c_void_p(h.H)


def _wrap(raw: C.c_void_p, csr_cls) -> cuda_h:
    nr, nc, nnz = C.c_int32(), C.c_int32(), C.c_int64()
    N.check(N.lib().csrk_dims(raw, C.byref(nr), C.byref(nc), C.byref(nnz), None, None), "dims")
    return cuda_h(raw.value, nr.value, nc.value, nnz.value, csr_cls)


def to_handle(csr) -> cuda_h:
    """Upload a CSR (csr/kernels/numba/__init__.py:16-27).  Copies: the matrix may
    be modified or dropped afterwards.  Accepts int32/int64 rowptrs, f4/f8 or
    absent values, and empty matrices."""
    nnz = int(csr.nnz)
    if nnz > max_nnz:
        raise ValueError("CSR size {} exceeds max nnz {}".format(nnz, max_nnz))
    rps = np.asarray(csr.rowptrs)
    if rps.dtype != np.int32 and rps.dtype != np.int64:
        rps = rps.astype(np.int64 if nnz > _INTC_MAX else np.int32)
    rps = np.ascontiguousarray(rps)
    cis = np.ascontiguousarray(np.asarray(csr.colinds)[:nnz], dtype=np.int32)
    vs = csr.values
    if vs is not None:
        vs = np.asarray(vs)[:nnz]
        if vs.dtype != np.float32 and vs.dtype != np.float64:
            vs = vs.astype(np.float64)
        vs = np.ascontiguousarray(vs)
    out = C.c_void_p()
    rc = N.lib().csrk_create(int(csr.nrows), int(csr.ncols), nnz, _ptr(rps), int(rps.dtype.itemsize == 8),
                             _ptr(cis), _ptr(vs), 0 if vs is None else vs.dtype.itemsize, C.byref(out))
    N.check(rc, "to_handle")
    return cuda_h(out.value, csr.nrows, csr.ncols, nnz, type(csr))


def from_handle(h: cuda_h):
    """Copy a handle back to a host CSR (numba/__init__.py:30-36).  The result is
    independent of the handle, which may be released right after."""
    raw = _live(h)
    nr, nc, nnz, is64, vk = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int(), C.c_int()
    L = N.lib()
    N.check(L.csrk_dims(raw, C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(is64), C.byref(vk)), "from_handle")
    rps = np.empty(nr.value + 1, np.int64 if is64.value else np.int32)
    cis = np.empty(nnz.value, np.int32)
    vs = None if vk.value == 0 else np.empty(nnz.value, np.float32 if vk.value == 4 else np.float64)
    N.check(L.csrk_export(raw, _ptr(rps), _ptr(cis), _ptr(vs)), "from_handle")
    cls = h.csr_cls or _default_csr_cls()
    try:
        # keep the handle's dtypes exactly (int64 rowptrs stay int64, as the numba kernel's do)
        return cls(nr.value, nc.value, nnz.value, rps, cis, vs, _cast=False)
    except TypeError:
        return cls(nr.value, nc.value, nnz.value, rps, cis, vs)


def release_handle(h: cuda_h) -> None:
    """Free the device memory (numba/__init__.py:39-44).  Like mkl_h, ``H`` is
    zeroed, so a second release is a no-op (csr/kernels/mkl/handle.py:144-148)."""
    if h.H:
        raw, h.H = C.c_void_p(h.H), 0
        N.check(N.lib().csrk_free(raw), "release_handle")


def order_columns(h: cuda_h) -> None:
    """Sort every row by column, in place, values carried along
    (numba/__init__.py:47-52 -> csr/structure.py:156-169)."""
    N.check(N.lib().csrk_order_columns(_live(h)), "order_columns")


def mult_ab(a_h: cuda_h, b_h: cuda_h) -> cuda_h:
    """C = A B as a new handle the caller releases (multiply.py:13-38)."""
    assert a_h.ncols == b_h.nrows
    out = C.c_void_p()
    N.check(N.lib().csrk_spgemm(_live(a_h), _live(b_h), C.byref(out)), "mult_ab")
    return _wrap(out, a_h.csr_cls)


def mult_abt(a_h: cuda_h, b_h: cuda_h) -> cuda_h:
    """C = A B^T as a new handle the caller releases (multiply.py:41-57)."""
    assert a_h.ncols == b_h.ncols
    out = C.c_void_p()
    N.check(N.lib().csrk_spgemm_abt(_live(a_h), _live(b_h), C.byref(out)), "mult_abt")
    return _wrap(out, a_h.csr_cls)


def mult_vec(h: cuda_h, v, out=None) -> np.ndarray:
    """y = A v as a host float64 vector (numba/__init__.py:55-67).  ``v`` may have
    any numeric dtype; float32 stays float32 on the device, everything else is
    promoted to float64.  ``out`` (optional, float64[nrows], e.g. pinned memory)
    receives the result."""
    raw = _live(h)
    x = np.asarray(v)
    if x.dtype != np.float32 and x.dtype != np.float64:
        x = x.astype(np.float64)
    x = np.ascontiguousarray(x)
    if x.shape != (h.ncols,):
        raise ValueError(f"vector has shape {x.shape}, expected ({h.ncols},)")
    if out is None:
        out = np.empty(h.nrows, np.float64)
    elif out.dtype != np.float64 or out.shape != (h.nrows,) or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of length nrows")
    N.check(N.lib().csrk_spmv(raw, _ptr(x), x.dtype.itemsize, _ptr(out)), "mult_vec")
    return out


# ---------------------------------------------------------------- extras

def transpose(h: cuda_h, include_values: bool = True) -> cuda_h:
    """Stable CSR->CSC as a new handle (csr/structure.py:172-247)."""
    out = C.c_void_p()
    N.check(N.lib().csrk_transpose(_live(h), int(bool(include_values)), C.byref(out)), "transpose")
    return _wrap(out, h.csr_cls)


def filter_zeros(h: cuda_h) -> None:
    """Drop stored zeros in place on the device (csr/_struct.py:61-79)."""
    raw = _live(h)
    N.check(N.lib().csrk_filter_zeros(raw), "filter_zeros")
    nnz = C.c_int64()
    N.check(N.lib().csrk_dims(raw, None, None, C.byref(nnz), None, None), "dims")
    h.nnz = nnz.value


def from_coo(rows, cols, vals, shape=None, csr_cls=None) -> cuda_h:
    """COO triples -> a resident handle, sorted by row on the device with the COO order kept inside
    each row (csr/csr.py:140-169 -> csr/structure.py:11-58).  ``from_handle`` gives the host CSR."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    assert np.min(rows, initial=0) >= 0
    assert np.min(cols, initial=0) >= 0
    if shape is not None:
        nrows, ncols = (int(v) for v in shape)
        assert np.max(rows, initial=0) < max(nrows, 1)
        assert np.max(cols, initial=0) < max(ncols, 1)
    else:
        nrows = int(np.max(rows)) + 1
        ncols = int(np.max(cols)) + 1
    nnz = len(rows)
    assert len(cols) == nnz
    assert vals is None or len(vals) == nnz
    if vals is not None:
        vals = np.ascontiguousarray(vals)
        if vals.dtype != np.float32:
            vals = vals.astype(np.float64, copy=False)
    out = C.c_void_p()
    N.check(N.lib().csrk_from_coo(nrows, ncols, nnz, _ptr(rows), _ptr(cols), _ptr(vals),
                                  0 if vals is None else vals.dtype.itemsize, C.byref(out)), "from_coo")
    return _wrap(out, csr_cls)


def normalize_rows(h: cuda_h, normalization: str, values_out=None) -> np.ndarray:
    """Normalise the rows of the handle's matrix in place on the device and return the per-row
    means ('center') or norms ('unit') in the values' dtype (csr/transform.py:13-66).
    ``values_out`` (optional, C-contiguous array of nnz elements of the values' dtype) receives the
    normalised values."""
    kinds = {"center": 0, "unit": 1}
    if normalization not in kinds:
        raise ValueError('unknown normalization: ' + normalization)
    raw = _live(h)
    vk = C.c_int()
    N.check(N.lib().csrk_dims(raw, None, None, None, None, C.byref(vk)), "dims")
    if vk.value not in (4, 8):
        raise ValueError("normalize_rows needs a matrix with values")
    dt = np.float32 if vk.value == 4 else np.float64
    vec = np.zeros(h.nrows, dt)
    if values_out is not None and (values_out.dtype != dt or values_out.shape != (h.nnz,)
                                   or not values_out.flags.c_contiguous):
        raise ValueError("values_out must be a C-contiguous array of nnz elements of the matrix's value type")
    N.check(N.lib().csrk_normalize_rows(raw, kinds[normalization], _ptr(vec), _ptr(values_out)), "normalize_rows")
    return vec


def subset_rows(h: cuda_h, begin: int, end: int) -> cuda_h:
    """Rows [begin, end) as a new handle (csr/structure.py:70-81), copied on the device."""
    out = C.c_void_p()
    N.check(N.lib().csrk_subset_rows(_live(h), int(begin), int(end), C.byref(out)), "subset_rows")
    return _wrap(out, h.csr_cls)


def spgemm_stats(h: cuda_h) -> dict:
    """Products P and out-nnz Z of the multiplication that produced ``h``."""
    p, z = C.c_int64(), C.c_int64()
    N.check(N.lib().csrk_spgemm_stats(_live(h), C.byref(p), C.byref(z)), "spgemm_stats")
    path = C.c_int()
    N.check(N.lib().csrk_spgemm_path(_live(h), C.byref(path)), "spgemm_path")
    side = C.c_int64()
    N.check(N.lib().csrk_spgemm_side_list(_live(h), C.byref(side)), "spgemm_side_list")
    return {"products": p.value, "out_nnz": z.value, "side_list": side.value,
            "dense_path": {0: "none", 1: "owner", 2: "fixed"}.get(path.value, str(path.value))}


def mult_vec_dev(h: cuda_h, x_ptr: int, x_itemsize: int, y_ptr: int, stream: int = 0) -> None:
    """Device-pointer SpMV, enqueued on the cudaStream_t ``stream`` (used verbatim: 0 is
    CUDA's default stream, e.g. ``torch.cuda.current_stream().cuda_stream``), no
    synchronisation: for callers that own device buffers (multi-GPU layer, bench)."""
    N.check(N.lib().csrk_spmv_dev(_live(h), C.c_void_p(x_ptr), int(x_itemsize), C.c_void_p(y_ptr),
                                  C.c_void_p(stream)), "mult_vec_dev")


def mult_vec_dev_multi(h: cuda_h, x_ptr: int, x_itemsize: int, y_ptrs, stream: int = 0) -> None:
    """Fused SpMV + gather: like :func:`mult_vec_dev`, but every finished row is stored to all
    of ``y_ptrs`` (``y_ptrs[0]`` local, the rest the same segment in peer GPUs' buffers)."""
    arr = (C.c_void_p * len(y_ptrs))(*[C.c_void_p(int(p)) for p in y_ptrs])
    N.check(N.lib().csrk_spmv_dev_multi(_live(h), C.c_void_p(x_ptr), int(x_itemsize), arr, len(y_ptrs),
                                        C.c_void_p(stream)), "mult_vec_dev_multi")


def mult_vec_dev_mc(h: cuda_h, x_ptr: int, x_itemsize: int, y_ptr: int, y_mc_ptr: int, stream: int = 0) -> None:
    """SpMV whose finished rows are also stored through the NVLink multicast address ``y_mc_ptr``
    (this rank's segment of a symmetric gather buffer): one store reaches every GPU."""
    N.check(N.lib().csrk_spmv_dev_mc(_live(h), C.c_void_p(x_ptr), int(x_itemsize), C.c_void_p(y_ptr),
                                     C.c_void_p(y_mc_ptr), C.c_void_p(stream)), "mult_vec_dev_mc")


def mc_broadcast(mc_dst_ptr: int, src_ptr: int, nbytes: int, stream: int = 0) -> None:
    "Copy local device memory to an NVLink multicast address (the root's side of the x broadcast)."
    N.check(N.lib().csrk_mc_broadcast(C.c_void_p(mc_dst_ptr), C.c_void_p(src_ptr), int(nbytes),
                                      C.c_void_p(stream)), "mc_broadcast")


def from_device_arrays(nrows, ncols, nnz, rowptrs_ptr, rp_is64, colinds_ptr, values_ptr, val_kind,
                       stream: int = 0, csr_cls=None) -> cuda_h:
    """Build a handle from device arrays (D2D copy)."""
    out = C.c_void_p()
    rc = N.lib().csrk_create_dev(int(nrows), int(ncols), int(nnz), C.c_void_p(rowptrs_ptr), int(rp_is64),
                                 C.c_void_p(colinds_ptr), C.c_void_p(values_ptr) if values_ptr else None,
                                 int(val_kind), C.c_void_p(stream), C.byref(out))
    N.check(rc, "from_device_arrays")
    return cuda_h(out.value, nrows, ncols, nnz, csr_cls)


def spmv_plan_info(h: cuda_h, x_itemsize: int = 4) -> dict:
    """Which SpMV kernel serves ``h`` for x of the given item size (the plan is built by the first
    ``mult_vec`` that selects it): ``{'kernel': 'stream' | 'tile', ...plan shape}``."""
    info = (C.c_int64 * 12)()
    N.check(N.lib().csrk_spmv_plan_info(_live(h), int(x_itemsize), info), "spmv_plan_info")
    names = ("ctas", "warps", "slabs", "slab_cols", "rows_per_warp", "pseudo_rows", "split_rows", "smem_bytes",
             "stream_bytes", "piece", "ring_bytes")
    out = {"kernel": "stream" if info[0] else "tile"}
    if info[0]:
        out.update({n: int(info[i + 1]) for i, n in enumerate(names)})
    return out


def set_option(name: str, value: int) -> None:
    """Library tunables (``include/csrk.h``): ``spmv_mode`` (0 auto, 1 CSR tile kernel, 2 slab-stream
    kernel), ``stream_min_nnz``, ``stream_slab_bytes``, ``stream_ctas``, ``stream_warps``, ..."""
    N.check(N.lib().csrk_set_option(name.encode(), int(value)), "set_option")


def library_stream() -> int:
    "The library's own stream as an integer cudaStream_t."
    out = C.c_void_p()
    N.check(N.lib().csrk_get_stream(C.byref(out)), "get_stream")
    return out.value or 0


def synchronize() -> None:
    N.check(N.lib().csrk_synchronize(), "synchronize")
