"""
CUDA kernel for B200 (sm_100a), provided by the ``csr_b200`` package.

This is the ONE file a lenskit/csr maintainer adds, as ``csr/kernels/cuda/__init__.py``
(INTEGRATION.md section 1).  The reference resolves kernels by module name
(``get_kernel('cuda')`` / ``CSR_KERNEL=cuda`` -> ``import_module('csr.kernels.cuda')``,
csr/kernels/__init__.py:81-97,100-116), and a kernel is a module with the eight attributes of
docs/kernels.rst:61-104.  Like ``csr.kernels.scipy`` it is never picked as the default.
"""
from csr_b200.kernels.cuda import (  # noqa: F401
    max_nnz, to_handle, from_handle, release_handle,
    order_columns, mult_ab, mult_abt, mult_vec,
    # extras beyond the contract (device transpose / zero filter / row slice)
    transpose, filter_zeros, subset_rows,
)
