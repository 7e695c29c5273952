"""
csr_b200 -- a B200-native (sm_100a) CUDA kernel backend for lenskit/csr.

Layout
    csr_b200.kernels        kernel selection API (get_kernel / set_kernel / use_kernel / releasing)
    csr_b200.kernels.cuda   the kernel module: to_handle / from_handle / release_handle /
                            order_columns / mult_ab / mult_abt / mult_vec (+ device extras)
    csr_b200.kernel         the default kernel bound statically (mirror of csr/kernel.py)
    csr_b200.CSR            host mirror of the reference CSR class (callers of the kernel)
    csr_b200.dist           row-partitioned multi-GPU SpMV / SpGEMM over torch.distributed (NCCL)
    csr_b200.synth          seeded synthetic workloads of BASELINE.json's configs
    csr_b200/csrc           hand-written CUDA kernels + the C ABI (include/csrk.h)
"""

from .csr import CSR  # noqa: F401
from .kernels import get_kernel, set_kernel, use_kernel, releasing  # noqa: F401

__version__ = "0.1.0"
