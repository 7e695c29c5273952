"""BASELINE configs[3] at full scale: transpose of the 5M x 5M / 500M-nnz matrix timed alone, and
mult_ab(A[r0:r1], A) on a row block (the full product does not fit in HBM: Z*12 B)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
t = time.perf_counter(); A = synth.cfg4_square(scale); print("gen", A, f"{time.perf_counter()-t:.1f}s", flush=True)
t = time.perf_counter(); ah = K.to_handle(A); print(f"to_handle {time.perf_counter()-t:.2f}s", flush=True)
for i in range(3):
    t = time.perf_counter(); th = K.transpose(ah); dt = time.perf_counter() - t
    b = A.nnz * 12 + (A.nrows + 1) * 4 + A.nnz * 12 + (A.ncols + 1) * 4
    print(f"transpose {dt*1e3:8.2f} ms  {A.nnz/dt/1e9:.2f} Gnnz/s  {b/dt/1e9:.1f} GB/s algorithmic", flush=True)
    if i < 2: K.release_handle(th)
# (A^T)^T == A structure check on the device result, via a second transpose and export of rowptrs only is costly; check nnz
tth = K.transpose(th); print("double transpose nnz", tth.nnz, "==", A.nnz); K.release_handle(tth); K.release_handle(th)
rows = int(A.nrows * float(sys.argv[2])) if len(sys.argv) > 2 else A.nrows // 100
bh = K.subset_rows(ah, 0, rows)
for i in range(2):
    t = time.perf_counter(); ch = K.mult_ab(bh, ah); dt = time.perf_counter() - t
    st = K.spgemm_stats(ch); K.release_handle(ch)
    print(f"mult_ab(A[0:{rows}], A) {dt*1e3:8.2f} ms  Z={st['out_nnz']} P={st['products']}  {st['out_nnz']/dt/1e9:.2f} Gnnz/s {st['products']/dt/1e9:.1f} Gprod/s", flush=True)
