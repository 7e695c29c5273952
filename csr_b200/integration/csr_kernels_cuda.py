"""
CUDA kernel for B200 (sm_100a), provided by the ``csr_b200`` package.

This is the ONE file a lenskit/csr maintainer adds, as ``csr/kernels/cuda/__init__.py``
(INTEGRATION.md section 1).  The reference resolves kernels by module name
(``get_kernel('cuda')`` / ``CSR_KERNEL=cuda`` -> ``import_module('csr.kernels.cuda')``,
csr/kernels/__init__.py:81-97,100-116), and a kernel is a module with the eight attributes of
docs/kernels.rst:61-104.  Like ``csr.kernels.scipy`` it is never picked as the default.

Nopython callers.  ``csr/kernel.py:9-16`` freezes the default kernel's functions and the overloads in
``csr/_wiring.py:116-151`` call them from ``@njit`` code (``CSR.multiply`` / ``CSR.mult_vec`` on a
``CSRType``), the way the MKL kernel serves them (csr/kernels/mkl/handle.py:77-93).  The functions below
therefore carry Numba overloads: in nopython code a handle is the library's RAW handle (an integer),
``to_handle`` copies the structref's arrays to the device, ``from_handle`` builds the reference's own
``CSR`` from the copied-out arrays (int32 rowptrs like the numba kernel's output, multiply.py:28).
"""
import numpy as np
from numba import types
from numba.extending import overload

from csr import CSR
from csr_b200.kernels import cuda_numba as _nb
from csr_b200.kernels.cuda import (  # noqa: F401
    max_nnz, to_handle, from_handle, release_handle,
    order_columns, mult_ab, mult_abt, mult_vec,
    # extras beyond the contract (device transpose / zero filter / row slice)
    transpose, filter_zeros, subset_rows,
)


def _is_raw(h):
    return isinstance(h, types.Integer)


@overload(to_handle)
def _to_handle_jit(csr):
    if not hasattr(csr, "has_values"):      # not a CSRType
        return None
    if csr.has_values:
        return lambda csr: _nb.create(csr.nrows, csr.ncols, csr.nnz, csr.rowptrs, csr.colinds, csr.values)
    return lambda csr: _nb.create_structure(csr.nrows, csr.ncols, csr.nnz, csr.rowptrs, csr.colinds)


@overload(from_handle)
def _from_handle_jit(h):
    if not _is_raw(h):
        return None

    def impl(h):
        nrows, ncols, nnz, rp, ci, vs = _nb.export_arrays(h)
        if nnz > 2147483647:
            raise OverflowError("from_handle: more than INT32_MAX entries")
        return CSR(nrows, ncols, nnz, rp.astype(np.intc), ci, vs)
    return impl


@overload(release_handle)
def _release_jit(h):
    if _is_raw(h):
        return lambda h: _nb.release_handle(h)


@overload(order_columns)
def _order_jit(h):
    if _is_raw(h):
        return lambda h: _nb.order_columns(h)


@overload(mult_ab)
def _mult_ab_jit(a_h, b_h):
    if _is_raw(a_h) and _is_raw(b_h):
        return lambda a_h, b_h: _nb.mult_ab(a_h, b_h)


@overload(mult_abt)
def _mult_abt_jit(a_h, b_h):
    if _is_raw(a_h) and _is_raw(b_h):
        return lambda a_h, b_h: _nb.mult_abt(a_h, b_h)


@overload(mult_vec)
def _mult_vec_jit(a_h, x):
    if _is_raw(a_h):
        return lambda a_h, x: _nb.mult_vec(a_h, x)
