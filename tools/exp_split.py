"""Experiment: how do the tile kernel and the slab kernel do on the heavy and light halves of cfg2?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csr_b200 import CSR, synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")

def bench(A, x, mode, reps=30):
    K.set_option("spmv_mode", mode)
    h = K.to_handle(A)
    xd = torch.from_numpy(x).cuda(); yd = torch.zeros(A.nrows, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    K.release_handle(h); K.set_option("spmv_mode", 0)
    b = A.nnz * 8 + (A.nrows + 1) * 4 + A.ncols * 4 + A.nrows * 8
    return ms, b / ms / 1e6

A = synth.cfg2_spmv()
x = synth.dense_vector(A.ncols, 77, "f4")
lens = np.diff(A.rowptrs)
TH = 567
def subset(mask):
    rows = np.flatnonzero(mask)
    l = lens[rows]; rp = np.zeros(len(rows) + 1, np.int64); np.cumsum(l, out=rp[1:])
    idx = np.repeat(A.rowptrs[rows].astype(np.int64) - rp[:-1], l) + np.arange(rp[-1])
    return CSR(len(rows), A.ncols, int(rp[-1]), rp, A.colinds[idx], A.values[idx])
H = subset(lens > TH); L = subset(lens <= TH)
print("heavy", H, "light", L, flush=True)
for name, M in (("full", A), ("heavy", H), ("light", L)):
    for mode, mn in ((1, "tile"), (3, "cell")):
        ms, gbs = bench(M, x, mode)
        print(f"{name:6s} {mn:5s} {ms*1000:8.1f} us  {gbs:8.1f} GB/s", flush=True)
