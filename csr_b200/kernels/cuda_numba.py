"""
Numba-callable entry points of the cuda kernel.

The reference lets nopython code reach its native kernels: the MKL shim's cffi module is registered
with Numba (csr/kernels/mkl/_api.py:8-19) and ``csr/kernel.py:9-16`` re-exports the active kernel's
functions for the ``@overload_method``s in ``csr/_wiring.py:116-151``.  This module is the same idea
for ``libcsr_cuda.so``: the C-ABI functions are bound a second time with integer-typed pointer
arguments (Numba calls ctypes functions natively; ``arr.ctypes.data`` is an integer there), and thin
``@njit(nogil=True)`` wrappers give nopython callers ``mult_vec / mult_ab / mult_abt /
release_handle / dims / export_arrays`` on RAW handles (``cuda_h.H``, an integer).

    h = K.to_handle(A)                       # object mode
    @njit
    def power_step(hh, x):                   # nopython
        y = cuda_numba.mult_vec(hh, x)
        return y / np.sqrt((y * y).sum())
    power_step(h.H, x)

Errors surface as exceptions raised from nopython code (``ValueError`` for a bad handle or shape,
``RuntimeError`` otherwise); ``csr_b200._native.last_error()`` has the message.
"""

import ctypes as C

import numpy as np
from numba import njit, types
from numba.extending import overload

from .. import _native as N

N.lib()  # fail loudly here if the library is missing (no CPU fallback)
_L = C.CDLL(N.LIB_PATH)  # a second binding of the same loaded library, with integer pointer types
_P = C.c_size_t


def _bind(name, *argtypes):
    f = getattr(_L, name)
    f.restype = C.c_int
    f.argtypes = list(argtypes)
    return f


_create = _bind("csrk_create", C.c_int32, C.c_int32, C.c_int64, _P, C.c_int, _P, _P, C.c_int, _P)
_dims = _bind("csrk_dims", _P, _P, _P, _P, _P, _P)
_export = _bind("csrk_export", _P, _P, _P, _P)
_spmv = _bind("csrk_spmv", _P, _P, C.c_int, _P)
_spgemm = _bind("csrk_spgemm", _P, _P, _P)
_spgemm_abt = _bind("csrk_spgemm_abt", _P, _P, _P)
_free = _bind("csrk_free", _P)
_order = _bind("csrk_order_columns", _P)


@njit(nogil=True)
def dims(h):
    "(nrows, ncols, nnz, rowptrs_are_int64, value_bytes) of a raw handle."
    a32 = np.zeros(2, np.int32)
    a64 = np.zeros(1, np.int64)
    ai = np.zeros(2, np.intc)
    rc = _dims(h, a32.ctypes.data, a32.ctypes.data + 4, a64.ctypes.data, ai.ctypes.data, ai.ctypes.data + 4)
    if rc != 0:
        raise ValueError("invalid cuda kernel handle")
    return a32[0], a32[1], a64[0], ai[0], ai[1]


@njit(nogil=True)
def _create_checked(nrows, ncols, nnz, rp_ptr, rp_is64, ci_ptr, vs_ptr, val_kind):
    out = np.zeros(1, np.uintp)
    rc = _create(nrows, ncols, nnz, rp_ptr, rp_is64, ci_ptr, vs_ptr, val_kind, out.ctypes.data)
    if rc == 1:
        raise ValueError("to_handle: bad argument")
    if rc == 2:
        raise MemoryError("to_handle: out of device memory")
    if rc != 0:
        raise RuntimeError("to_handle failed")
    return out[0]


@njit(nogil=True)
def create(nrows, ncols, nnz, rowptrs, colinds, values):
    """to_handle from nopython code (numba/__init__.py:16-27): the six fields of the CSR record
    (csr/_struct.py:10-28) are copied to the device; returns a RAW handle the caller releases.  The copies
    are complete when this returns, so the arrays need not outlive the call."""
    rp = np.ascontiguousarray(rowptrs)
    ci = np.ascontiguousarray(colinds)
    vs = np.ascontiguousarray(values)
    h = _create_checked(nrows, ncols, nnz, rp.ctypes.data, 1 if rp.itemsize == 8 else 0, ci.ctypes.data, vs.ctypes.data,
                        vs.itemsize)
    if rp.shape[0] + ci.shape[0] + vs.shape[0] < 0:   # keeps rp / ci / vs alive across the native call
        h = np.uintp(0)
    return h


@njit(nogil=True)
def create_structure(nrows, ncols, nnz, rowptrs, colinds):
    "``create`` for a CSR without values (values is None: the kernels use 1)."
    rp = np.ascontiguousarray(rowptrs)
    ci = np.ascontiguousarray(colinds)
    h = _create_checked(nrows, ncols, nnz, rp.ctypes.data, 1 if rp.itemsize == 8 else 0, ci.ctypes.data, 0, 0)
    if rp.shape[0] + ci.shape[0] < 0:
        h = np.uintp(0)
    return h


def _kernel_vector(x):  # pragma: no cover - replaced by the overload below in nopython code
    raise NotImplementedError


@overload(_kernel_vector)
def _ov_kernel_vector(x):
    # float32 stays float32 on the device, every other dtype is promoted to float64 (kernels/cuda.py mult_vec)
    if isinstance(x, types.Array) and x.dtype in (types.float32, types.float64):
        return lambda x: np.ascontiguousarray(x)
    return lambda x: np.ascontiguousarray(x).astype(np.float64)


@njit(nogil=True)
def mult_vec(h, x):
    "y = A x as a host float64 vector (numba/__init__.py:55-67) for a raw handle ``h``."
    nrows, ncols, nnz, is64, vk = dims(h)
    if x.ndim != 1 or x.shape[0] != ncols:
        raise ValueError("vector length does not match the matrix")
    xv = _kernel_vector(x)
    y = np.empty(nrows, np.float64)
    rc = _spmv(h, xv.ctypes.data, xv.itemsize, y.ctypes.data)
    # xv may be a temporary (dtype promotion / contiguous copy).  Numba releases a variable after its LAST
    # use, and taking ``xv.ctypes.data`` is a use that ends before the call: without this read the buffer
    # could be freed (and reused by another nogil thread) while the native code still reads it.
    if xv.shape[0] != ncols:
        rc = 1
    if rc == 1:
        raise ValueError("mult_vec: bad argument")
    if rc != 0:
        raise RuntimeError("mult_vec failed")
    return y


@njit(nogil=True)
def _product(ah, bh, transpose):
    out = np.zeros(1, np.uintp)
    rc = _spgemm_abt(ah, bh, out.ctypes.data) if transpose else _spgemm(ah, bh, out.ctypes.data)
    if rc == 1:
        raise ValueError("shape mismatch or invalid handle")
    if rc == 2:
        raise MemoryError("mult_ab: out of device memory")
    if rc != 0:
        raise RuntimeError("mult_ab failed")
    return out[0]


@njit(nogil=True)
def mult_ab(ah, bh):
    "C = A B as a NEW raw handle the caller releases (multiply.py:13-38)."
    return _product(ah, bh, False)


@njit(nogil=True)
def mult_abt(ah, bh):
    "C = A B^T as a NEW raw handle the caller releases (multiply.py:41-57)."
    return _product(ah, bh, True)


@njit(nogil=True)
def order_columns(h):
    "Sort every row by column in place (numba/__init__.py:47-52)."
    if _order(h) != 0:
        raise RuntimeError("order_columns failed")


@njit(nogil=True)
def release_handle(h):
    "Free a raw handle (numba/__init__.py:39-44).  Releasing 0 is a no-op."
    if h != 0 and _free(h) != 0:
        raise RuntimeError("release_handle failed")


@njit(nogil=True)
def export_arrays(h):
    """(nrows, ncols, nnz, rowptrs:int64, colinds:int32, values:float64) copied from a raw handle
    (the nopython counterpart of from_handle; float32 values are widened, a structure-only matrix
    gives an empty values array)."""
    nrows, ncols, nnz, is64, vk = dims(h)
    ci = np.empty(nnz, np.int32)
    if is64:
        rp = np.empty(nrows + 1, np.int64)
        rpp = rp.ctypes.data
    else:
        rp32 = np.empty(nrows + 1, np.int32)
        rpp = rp32.ctypes.data
    if vk == 4:
        v4 = np.empty(nnz, np.float32)
        vp = v4.ctypes.data
    else:
        v8 = np.empty(nnz if vk == 8 else 0, np.float64)
        vp = v8.ctypes.data
    if _export(h, rpp, ci.ctypes.data, vp) != 0:
        raise RuntimeError("export failed")
    if not is64:
        rp = rp32.astype(np.int64)
    if vk == 4:
        v8 = v4.astype(np.float64)
    return nrows, ncols, nnz, rp, ci, v8
