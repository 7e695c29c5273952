"""Feasibility: time the tile SpMV on the light rows / heavy rows of cfg2 separately."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csr_b200 import synth
from csr_b200.csr import CSR
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
A = synth.cfg2_spmv(scale); x = synth.dense_vector(A.ncols, 77, "f4")
xd = torch.from_numpy(x).cuda(); yd = torch.zeros(A.nrows, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
lens = np.diff(A.rowptrs)

def timeit(M, label):
    h = K.to_handle(M)
    for _ in range(5): K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{label:28s} rows_nonempty={int((np.diff(M.rowptrs)>0).sum()):8d} nnz={M.nnz:10d}  {ms:.4f} ms  {M.nnz/ms/1e6:.1f} Gnnz/s", flush=True)
    K.release_handle(h)

timeit(A, "all")
for T in (64, 256, 1024):
    for name, keep_rows in (("light", lens < T), ("heavy", lens >= T)):
        keep = np.repeat(keep_rows, lens)
        l2 = np.where(keep_rows, lens, 0)
        rp = np.zeros(A.nrows + 1, np.int64); np.cumsum(l2, out=rp[1:])
        M = CSR(A.nrows, A.ncols, int(rp[-1]), rp.astype(np.int32), A.colinds[keep], A.values[keep])
        timeit(M, f"{name} T={T}")
